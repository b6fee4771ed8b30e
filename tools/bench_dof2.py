"""BASELINE config 3 (2-DOF, 4x4-block kernel): N = 8192 training pairs (n = 32768), 1e5 orbits -- training NLL+gradient
and map prediction throughput.   python tools/bench_dof2.py [N] [E] [steps] > gpurun_out/dof2.json"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as O
from sympgpr_b200 import api

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
E = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
x, z = O.henon_like_training(N)
for shrink in (1.0, 0.8, 0.65, 0.5):          # the first length scale the fp64 Cholesky accepts
    hyp = np.array([0.35 * shrink * (200.0 / N) ** 0.25, 0.4 * shrink * (200.0 / N) ** 0.25, 2 * np.max(np.abs(z)) ** 2, 1e-6])
    try:
        api.nll_grad4(hyp, x, z, 4 * N)
        break
    except np.linalg.LinAlgError:
        continue
t0 = time.perf_counter(); v, g = api.nll_grad4(hyp, x, z, 4 * N); t_nll = time.perf_counter() - t0
f = api.fit(hyp, x, z, 4 * N, reg=4)
q0 = np.vstack((-0.3 + 0.6 * O.halton(E, 2, start=7), -0.3 + 0.6 * O.halton(E, 3, start=7)))
p0 = np.vstack((-0.3 + 0.6 * O.halton(E, 5, start=7), -0.3 + 0.6 * O.halton(E, 7, start=7)))
api.applymap4(3, E, hyp[:3], q0, p0, x, f["alpha"], out_every=0)
t0 = time.perf_counter()
qf, pf, st = api.applymap4(steps + 1, E, hyp[:3], q0, p0, x, f["alpha"], out_every=0, return_stats=True)
t_map = time.perf_counter() - t0
dt = 0.3
P = np.vstack((p0[0] - dt * (q0[0] + 2 * q0[0] * q0[1]), p0[1] - dt * (q0[1] + q0[0] ** 2 - q0[1] ** 2)))
q1, p1 = api.applymap4(2, E, hyp[:3], q0, p0, x, f["alpha"], out_every=0)
print(json.dumps({"config": "03: 2-DOF 4x4-block kernel", "N": N, "n": 4 * N, "hyp": list(map(float, hyp)),
                  "nll_grad_s": t_nll, "nll_grad_tflops": (4.0 * N) ** 3 / t_nll / 1e12, "nll": v, "grad": list(map(float, g)),
                  "orbits": E, "steps": steps, "map_s": t_map, "orbit_steps_per_s": E * steps / t_map,
                  "passes_per_orbit_step": st["evaluations"] / (E * steps), "unconverged": st["unconverged"],
                  "pair_evals_per_s": st["evaluations"] * N / t_map,
                  "one_step_error_vs_training_map": float(max(np.abs(p1 - P).max(), np.abs(q1 - (q0 + dt * P)).max())),
                  "finite_final": float(np.isfinite(qf).mean())}))
