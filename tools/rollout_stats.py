"""1000-step rollout statistics, CUDA vs oracle (prints the series used by test_rollout_statistics_1000_steps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gmu_water_simulation_b200 as gws
from oracle_binding import Oracle
box = float(sys.argv[1]) if len(sys.argv) > 1 else 0.4
o = Oracle(box).setup_scene(); sim = gws.Simulator("cuda", box).setup_scene(); ctx = sim.context()
rows = []
for k in range(100):
    o.step(10); sim.step_many(10)
    so, sg = o.stats(), ctx.stats()
    rows.append((10 * (k + 1), so["ke"], sg["ke"], np.abs(so["com"] - sg["com"]).max(), so["fill"], sg["fill"]))
rows = np.array(rows)
np.set_printoptions(precision=5, suppress=True, linewidth=200)
print("step  ke_oracle  ke_cuda  com_err  fill_oracle  fill_cuda")
print(rows[::5])
print("max |dKE|/max KE", np.abs(rows[:, 1] - rows[:, 2]).max() / rows[:, 1].max(), "max com err / h", rows[:, 3].max() / 0.0457,
      "max fill err / h", np.abs(rows[:, 4] - rows[:, 5]).max() / 0.0457)
rel = np.abs(rows[:, 2] / rows[:, 1] - 1)
print("rel KE err first 30 samples max", rel[:30].max(), "all", rel.max())
