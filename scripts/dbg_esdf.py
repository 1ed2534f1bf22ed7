import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200 import workloads, capi
from test_esdf_gpu import make_sdf
import oracle_lib
ctx = alore.Context(0)
for shape in [(64, 64), (20, 90)]:
    glx, gly = shape
    grid = workloads.random_map(glx, gly, seed=glx * 7 + gly, p_occ=0.0, p_unknown=0.05, wall=False)
    m = make_sdf(ctx, glx, gly, 0.05, grid)
    m.updateESDF2d()
    pos, neg = m.last_squared()
    print(shape, "pos uniq", np.unique(pos)[:5], "neg uniq", np.unique(neg)[:5], "dist[0]", m.distance_buffer_all_[0])
