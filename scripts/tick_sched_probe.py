"""Does the run-time schedule prediction carry across perturbed ticks?  (developer probe)"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import alore_legged_manipulator_b200 as alore
from alore_legged_manipulator_b200.ms_planner import DeviceBatch
import bench
ctx = alore.Context(0)
prm = alore.default_params()
geom, grid0 = bench.build_world()
m = alore.SDFmap(ctx, gridmap_interval=0.05, detection_range=1e6, global_x_lower=geom.x_lower,
                 global_x_upper=geom.x_lower + (geom.glx - 0.5) * 0.05, global_y_lower=geom.y_lower,
                 global_y_upper=geom.y_lower + (geom.gly - 0.5) * 0.05)
m.gridmap_[:] = grid0
m.has_map_ = True
m.forceUpdateESDF()
gm = m.geom()
pts0 = bench.way_points(gm, grid0, m.distance_buffer_all_.copy())
ticks = []
for t in range(5):
    g, p = bench.tick_variant(t, gm, grid0, pts0)
    ticks.append((g, bench.candidates_from_points(p, 0, 8320)))
batches = [DeviceBatch(ctx, c) for _, c in ticks]
ev = []
for rnd in range(2):
    for t in range(5):
        m.gridmap_[:] = ticks[t][0]
        m.updateESDF2d()
        batches[t].run(prm)
        r = batches[t].download()
        ev.append(r.evals.copy())
        print(f"round {rnd} tick {t}: kernel ms {batches[t].kernel_ms():.1f}  evals identical to previous run: "
              f"{(ev[-1] == ev[-2]).mean() if len(ev) > 1 else float('nan'):.3f}", flush=True)
