"""Multi-GPU frames: one process per GPU (torch.distributed), full scene replica per GPU, the ray
planes of every frame dealt out in interleaved blocks, finished pixels composited on rank 0 with one
reduce over NVLink (NCCL).  Nothing in the reference corresponds to this (it is single-GPU,
SURVEY.md §2a); the split follows SURVEY.md §8e: ray planes are independent, the only exchange is the
final compositing, and the multi-GPU frame must equal the single-GPU frame bit for bit.

Slices are interleaved rather than contiguous because ray cost varies smoothly with the ray index
(rays toward the horizon walk ~20x more cells than rays toward the ground): dealing blocks of
`block` consecutive ray planes round-robin gives every GPU the same mix.
"""
import ctypes as C

DEFAULT_BLOCK = 32


def owned_mask(count, block, nranks, rank):
    """Boolean list: does `rank` own ray plane r (r in [0, count))?"""
    return [((r // block) % nranks) == rank for r in range(count)]


def owned_count(count, block, nranks, rank):
    if nranks <= 1:
        return count
    cyc = block * nranks
    owned = (count // cyc) * block
    rem = count % cyc - rank * block
    return owned + max(0, min(rem, block))


def composite(image, dist, dst=0):
    """Sum-reduce the per-rank images (uint8 tensors with disjoint support) onto rank `dst`.
    NCCL over NVLink on GPUs; gloo in the CPU tests.  Returns `image` (complete on rank dst)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(image, dst=dst, op=dist.ReduceOp.SUM)
    return image


class SlicedFrame:
    """Per-rank driver of a sliced frame: traverse own ray planes -> unwarp own pixels -> composite."""

    def __init__(self, renderer, cfg, torch, rank=0, world=1, dist=None, block=DEFAULT_BLOCK):
        self.r, self.cfg, self.dist, self.torch, self.block = renderer, cfg, dist, torch, block
        self.rank, self.world = rank, world
        dev = torch.device("cuda", renderer.device)
        self.rgba = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device=dev)
        # kernels and the collective share torch's current stream: ordering without host syncs
        renderer.set_stream(torch.cuda.current_stream(dev).cuda_stream)

    def render(self, raymap_gpu):
        if self.world == 1:
            self.r.render(raymap_gpu, self.cfg)
            self.r.unwarp(raymap_gpu, self.cfg, d_rgba=self.rgba.data_ptr())
            return self.rgba
        self.r.render_interleaved(raymap_gpu, self.cfg, self.block, self.world, self.rank)
        self.r.unwarp_interleaved(raymap_gpu, self.cfg, self.block, self.world, self.rank, d_rgba=self.rgba.data_ptr())
        return composite(self.rgba, self.dist)


class FrameFarm:
    """Alternate-frame rendering for throughput: the frames of a sequence are dealt round-robin to the
    ranks (rank r renders frames r, r+N, ...), every rank renders WHOLE frames from its own replica,
    and the finished frames of a round are gathered on rank 0 with one NCCL gather that overlaps the
    next round's rendering (double-buffered).  This is BASELINE config 5's "one camera per GPU with
    NVLink compositing" applied to the fly-through; per-frame latency is that of one GPU."""

    def __init__(self, renderer, cfg, torch, rank, world, dist, device=None):
        self.r, self.cfg, self.torch, self.rank, self.world, self.dist = renderer, cfg, torch, rank, world, dist
        # device: only the CPU tests of the dealing / gathering logic (gloo, a stand-in renderer) pass one
        dev = device if device is not None else torch.device("cuda", renderer.device)
        shape = (cfg.height, cfg.width, 4)
        self.bufs = [torch.zeros(shape, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.lists = [[torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
                      for _ in range(2)]
        self.handles = [None, None]
        if dev.type == "cuda":
            renderer.set_stream(torch.cuda.current_stream(dev).cuda_stream)

    def render_round(self, rd, raymap_gpu):
        """Render this rank's frame of round `rd` (None: no frame left) and start gathering the round."""
        k = rd & 1
        if self.handles[k] is not None:
            self.handles[k].wait()          # buffer k is free again (stream-ordered, no host block)
        if raymap_gpu is not None:
            self.r.render(raymap_gpu, self.cfg)
            self.r.unwarp(raymap_gpu, self.cfg, d_rgba=self.bufs[k].data_ptr())
        self.handles[k] = self.dist.gather(self.bufs[k], self.lists[k], dst=0, async_op=True)

    def wait_round(self, rd):
        k = rd & 1
        if self.handles[k] is not None:
            self.handles[k].wait()
            self.handles[k] = None
        return self.lists[k]

    def finish(self):
        for k in (0, 1):
            if self.handles[k] is not None:
                self.handles[k].wait()
                self.handles[k] = None
