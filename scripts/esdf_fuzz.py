"""Sequential fuzz of the ESDF path on ONE context (stale-state exposure): random patchwork maps, sizes around the tile /
band / vector boundaries, random windows, distances and squared planes against the CPU oracle, bit for bit.
usage: python scripts/esdf_fuzz.py [cases] [seed]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from test_esdf_gpu import check_against_oracle, patchwork_map
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
rng = np.random.default_rng(seed)
ctx = alore.Context(0)
edges = [16, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 191, 192, 255, 256, 257, 300, 383, 384, 385, 511, 512, 513, 640]
for it in range(n):
    glx = int(rng.choice(edges)) if rng.random() < 0.6 else int(rng.integers(8, 700))
    gly = int(rng.choice(edges)) if rng.random() < 0.6 else int(rng.integers(8, 700))
    grid = patchwork_map(glx, gly, rng)
    gi = 0.1
    if rng.random() < 0.4:
        m = check_against_oracle(ctx, glx, gly, gi, grid)
    else:
        odom = (float(rng.uniform(-0.45, 0.45) * glx * gi), float(rng.uniform(-0.45, 0.45) * gly * gi))
        m = check_against_oracle(ctx, glx, gly, gi, grid, odom=odom, rng_m=float(rng.uniform(1.0, 0.5 * max(glx, gly) * gi)))
    m.close()
print(f"esdf fuzz: {n} cases ok (seed {seed})")
