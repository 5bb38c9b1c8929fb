"""Randomised parity run: production kernel vs the oracle port on random cameras / scenes / resolutions / flags.
(tests/, __graft_entry__.smoke() and bench.py are the places allowed to use oracle/; this tool is a test helper run by
hand: `python tools/fuzz_parity.py N [SEED]`, exit code 1 on the first mismatch.)"""
import importlib, math, os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
R = importlib.import_module("rle-based-voxel-raycasting_b200")
from oracle import refbind as rb
from util import oracle_raymap

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
scenes = [("terrain128", R.RLE4.synth(0, 128, 128, 128, seed=1), 128), ("runs128", R.RLE4.synth(1, 128, 128, 128, seed=42), 128),
          ("rle256", R.RLE4.synth_rle(256, 256, 256, seed=42, band_every=8), 256), ("terrain64", R.RLE4.synth(0, 64, 64, 64, seed=3), 64)]
from util import edge_scenes
scenes += [("edge_" + k, v, 64) for k, v in edge_scenes(R).items()]
r = R.Renderer(0)
r.set_lanes_per_ray(int(os.environ.get("FUZZ_LANES", "0")))      # 0 = automatic choice, 65 = k_traverse_f, 68 = k_traverse_p
bad = 0
for i in range(N):
    name, scene, size = rnd.choice(scenes)
    wh = rnd.choice([(320, 240), (640, 480), (1024, 768), (1920, 1080), (800, 200), (512, 512)])
    cfg = R.FrameConfig.default(*wh)
    cfg.flags = rnd.choice([0, 0, 0, 1, 2, 3])
    if rnd.random() < 0.3:
        cfg.z_far = rnd.choice([500, 3000, 20000])
    if rnd.random() < 0.3:
        cfg.mip_distance = rnd.choice([64, 200, 700])
    mode = rnd.random()
    if mode < 0.4:      # far outside, infinite tiling
        pos = (rnd.uniform(-20000, 20000), -rnd.uniform(0.05, 1.5) * size, rnd.uniform(-20000, 20000))
    elif mode < 0.8:    # above / inside the volume
        pos = (rnd.uniform(-size, 2 * size), -rnd.uniform(-0.8, 1.2) * size, rnd.uniform(-size, 2 * size))
    else:               # lattice-aligned
        pos = (float(rnd.randint(-300, 300)), -float(rnd.randint(5, size)), float(rnd.randint(-300, 300)))
    rot = (rnd.choice([rnd.uniform(-1.5, 1.5), rnd.uniform(-3.1, 3.1), 0.0, 0.35]), rnd.choice([rnd.uniform(0, 6.3), 0.0, math.pi / 2, math.pi]), 0.0)
    if os.environ.get("FUZZ_VERBOSE"):
        print("case", i, name, wh, cfg.flags, cfg.z_far, cfg.mip_distance, pos, rot, flush=True)
    r.all_to_gpu(scene)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    orm = oracle_raymap(rb, rm, scene)
    want, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, flags=cfg.flags)
    wp = r.warp_buffer(cfg)
    r.upload(wp, np.zeros((cfg.rays_casted, cfg.render_size), np.uint32))
    r.render(rm, cfg)
    got = r.read_warp(cfg)
    if not np.array_equal(got, want):
        bad += 1
        print("MISMATCH case %d: %s %s flags %d z_far %d mipd %d pos %s rot %s: %d texels differ" % (i, name, wh, cfg.flags, cfg.z_far, cfg.mip_distance, pos, rot, int((got != want).sum())), flush=True)
        if bad >= 3:
            break
print("fuzz: %d cases, %d mismatches" % (i + 1, bad))
sys.exit(1 if bad else 0)
