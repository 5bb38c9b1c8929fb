"""Multi-GPU frames: one process per GPU (torch.distributed), full scene replica per GPU, the ray
planes of every frame dealt out in interleaved blocks ("column slices" of the warped ray buffer), finished
pixels composited over NVLink with NCCL.  Nothing in the reference corresponds to this (it is single-GPU,
SURVEY.md §2a); the split follows SURVEY.md §8e: ray planes are independent (R/src/Cuda_Render.h:109: every ray
plane writes only its own row), the only exchange is the final compositing, and the multi-GPU frame must equal
the single-GPU frame bit for bit.

Slices are interleaved rather than contiguous because ray cost varies smoothly with the ray index
(rays toward the horizon walk ~20x more cells than rays toward the ground): dealing blocks of
`block` consecutive ray planes round-robin gives every GPU the same mix.

Two compositors, both exact (the per-rank images have disjoint support, so a sum is a select):
  reduce   one NCCL reduce(sum) of the RGBA8 images to rank 0: the finished frame is on GPU 0.
  bands    one NCCL reduce_scatter(sum): rank r ends up with rows [r*H/N, (r+1)*H/N) of the finished frame and
           copies them to the host itself, so a frame that is wanted in HOST memory leaves over N PCIe links.
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


def band_rows(height, nranks):
    """Rows per rank of the `bands` compositor (the image is padded to nranks * band_rows rows)."""
    return (height + nranks - 1) // nranks


def composite(image, dist, dst=0):
    """Sum-reduce the per-rank images (uint8 tensors with disjoint support) onto rank `dst`.
    NCCL over NVLink on GPUs; gloo in the CPU tests.  Returns `image` (complete on rank dst)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(image, dst=dst, op=dist.ReduceOp.SUM)
    return image


def composite_bands(image, band, dist):
    """reduce_scatter(sum) of the per-rank images: `band` (this rank's rows of the finished frame) <- sum over
    ranks of the matching rows of `image` ([nranks * band_rows, W, 4]).  gloo has no reduce_scatter: the CPU
    tests go through all_reduce + slicing, the same arithmetic."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    if dist.get_backend() == "nccl":
        return dist.reduce_scatter_tensor(band, image, op=dist.ReduceOp.SUM, async_op=True)
    dist.all_reduce(image, op=dist.ReduceOp.SUM)
    rows = image.shape[0] // world
    band.copy_(image[rank * rows:(rank + 1) * rows])
    return None


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


class FramePipe:
    """`depth` frames in flight on this rank's GPU.  One context per slot (its own stream, warped buffer, DDA states
    and RGBA image), all sharing ONE scene replica (rlerc_scene_share).  Frame i goes to slot i % depth; with
    world > 1 every rank renders its interleaved slice of every frame and the compositor runs on the slot's stream
    order (torch's NCCL stream waits for the slot's kernels; the slot waits for the collective before it is reused),
    so nothing blocks the host and the collectives are issued in the same order on every rank.

    compose: "reduce" (frame complete on rank 0), "bands" (reduce_scatter, every rank keeps its rows; `host` = a
    [slots, nranks * band_rows, W, 4] uint8 host tensor shared by all ranks gets every band copied in by its owner),
    "none" (world == 1, or frames kept where they were rendered)."""

    def __init__(self, R, torch, device, scene, cfg, depth=4, rank=0, world=1, dist=None, block=DEFAULT_BLOCK, lanes=0,
                 compose="reduce", host=None, share_from=None):
        self.torch, self.dist, self.cfg = torch, dist, cfg
        self.rank, self.world, self.block, self.depth = rank, world, block, depth
        self.compose = compose if world > 1 else "none"
        dev = torch.device("cuda", device)
        self.r = [R.Renderer(device) for _ in range(depth)]
        if share_from is not None:
            for x in self.r:
                x.share_scene(share_from)
        else:
            self.r[0].all_to_gpu(scene)
            for x in self.r[1:]:
                x.share_scene(self.r[0])
        self.streams = [torch.cuda.Stream(dev) for _ in range(depth)]
        for x, s in zip(self.r, self.streams):
            x.set_lanes_per_ray(lanes)
            x.set_stream(s.cuda_stream)
        self.rows = band_rows(cfg.height, world) if self.compose == "bands" else cfg.height
        full = self.rows * world if self.compose == "bands" else cfg.height
        self.rgba = [torch.zeros((full, cfg.width, 4), dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.band = [torch.zeros((self.rows, cfg.width, 4), dtype=torch.uint8, device=dev) for _ in range(depth)] \
            if self.compose == "bands" else None
        self.pending = [None] * depth
        self.host = host

    def submit(self, i, raymap_gpu):
        """Enqueue frame i (asynchronous).  Returns the slot it went to."""
        k = i % self.depth
        with self.torch.cuda.stream(self.streams[k]):
            if self.pending[k] is not None:
                self.pending[k].wait()              # stream-ordered: the slot's image is free again
                self.pending[k] = None
            self.r[k].frame_device(raymap_gpu, self.cfg, self.block, self.world, self.rank, self.rgba[k].data_ptr())
            if self.compose == "reduce":
                self.pending[k] = self.dist.reduce(self.rgba[k], dst=0, op=self.dist.ReduceOp.SUM, async_op=True)
            elif self.compose == "bands":
                w = composite_bands(self.rgba[k], self.band[k], self.dist)
                if w is not None:
                    w.wait()
                if self.host is not None:
                    lo = self.rank * self.rows
                    self.host[k, lo:lo + self.rows].copy_(self.band[k], non_blocking=True)
            elif self.host is not None:
                self.host[k].copy_(self.rgba[k], non_blocking=True)
        return k

    def finish_slot(self, k):
        """Make slot k's stream wait for its compositing collective."""
        if self.pending[k] is not None:
            with self.torch.cuda.stream(self.streams[k]):
                self.pending[k].wait()
            self.pending[k] = None

    def image_numpy(self, k):
        return self.image(k).cpu().numpy(), 0, self.cfg.height

    def drain(self, onto=None):
        """Make `onto` (default: the current stream) wait for everything submitted so far."""
        cur = onto if onto is not None else self.torch.cuda.current_stream()
        for k, s in enumerate(self.streams):
            if self.pending[k] is not None:
                with self.torch.cuda.stream(s):
                    self.pending[k].wait()
                self.pending[k] = None
            cur.wait_stream(s)

    def start_after(self, event):
        for s in self.streams:
            s.wait_event(event)

    def image(self, k):
        """Slot k's finished frame as the compositor left it on this rank (reduce: complete on rank 0; bands: this
        rank's rows; none: the whole frame)."""
        if self.compose == "bands":
            return self.band[k]
        return self.rgba[k][:self.cfg.height]

    def close(self):
        self.torch.cuda.synchronize()
        for x in reversed(self.r):
            x.close()


class GroupPipe:
    """`depth` frames in flight through the library's own multi-GPU path (csrc/group.cu, include/rlerc.h "multi-GPU"):
    every rank traverses its interleaved ray-plane slices and produces its band of window rows with the texels pulled
    from the owners' warped buffers over NVLink; flag barriers in peer memory; no collective moves pixel data.
    Same interface as FramePipe.  dst >= 0: the finished frame is assembled on rank dst; dst = -1: every rank keeps
    its band (and copies it into `host`, a [slots, H, W, 4] uint8 host tensor shared by all ranks, when given)."""

    def __init__(self, R, torch, device, scene, cfg, depth=4, rank=0, world=1, dist=None, block=DEFAULT_BLOCK, lanes=0,
                 dst=0, host=None, share_from=None, views=False):
        """views: BASELINE config 5 — every rank renders the WHOLE frame of its own camera and delivers it into rank
        dst's view array over NVLink (rlerc_group_submit_view) instead of a slice of a common frame."""
        import numpy as np
        self.views = views
        self.np, self.torch, self.dist, self.cfg = np, torch, dist, cfg
        self.rank, self.world, self.block, self.depth, self.dst, self.host = rank, world, block, depth, dst, host
        r = R.Renderer(device)
        if share_from is not None:
            r.share_scene(share_from)
        else:
            r.all_to_gpu(scene)
        r.set_lanes_per_ray(lanes)
        self.r = [r]
        self.g = R.Group(r, cfg, rank, world, depth=depth, block=block, views=views)
        if world > 1:
            self.g.connect_distributed(torch, dist)
        dev = torch.device("cuda", device)
        self.streams = [torch.cuda.ExternalStream(self.g.stream(k), device=dev) for k in range(depth)]
        self.pending = [None] * depth
        self.last_ticket = -1

    def submit(self, i, raymap_gpu):
        """Enqueue the next frame (asynchronous; `i` is ignored: frames go to the slots in turn). Returns its slot."""
        host = None
        if self.host is not None:
            host = self.host[(self.last_ticket + 1) % self.depth].data_ptr()
        if self.views:
            self.last_ticket = self.g.submit_view(raymap_gpu, self.dst, host)
        else:
            self.last_ticket = self.g.submit(raymap_gpu, self.dst, host)
        return self.last_ticket % self.depth

    def finish_slot(self, k):
        pass                                     # barriers and copies are stream-ordered inside the slot

    def drain(self, onto=None):
        cur = onto if onto is not None else self.torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)

    def start_after(self, event):
        for s in self.streams:
            s.wait_event(event)

    def image_numpy(self, k):
        """Slot k's [H][W][4] image on this rank as numpy, and the rows this rank produced."""
        ptr, a, b = self.g.image(k)
        return self.r[0].download(ptr, (self.cfg.height, self.cfg.width, 4), self.np.uint8), a, b

    def view_numpy(self, k, r):
        """View r of slot k as it arrived on this rank (views mode, rank dst)."""
        ptr, stride = self.g.views(k)
        return self.r[0].download(ptr + r * stride, (self.cfg.height, self.cfg.width, 4), self.np.uint8)

    def close(self):
        self.torch.cuda.synchronize()
        self.g.close()
        self.r[0].close()


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
