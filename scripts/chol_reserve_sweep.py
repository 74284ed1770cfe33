"""Developer sweep: SMs kept free of the persistent tcgen05 trailing updates (far / near) vs the fit time (env knobs read at first use, so one process per setting)."""
import os, subprocess, sys, json
for N, D in ((4096, 8), (8192, 16)):
    for far, near in ((40, 16), (56, 16), (64, 16), (72, 16), (40, 32), (64, 32), (56, 48)):
        env = dict(os.environ, B200BO_I8_RESERVE=str(far), B200BO_I8_NEAR_RESERVE=str(near))
        out = subprocess.run([sys.executable, "scripts/fit_once.py", str(N), str(D)], env=env, capture_output=True, text=True).stdout.strip().splitlines()
        best = min(json.loads(l)["chol_ms"] for l in out)
        print(f"N={N} far_reserve={far} near_reserve={near}: chol {best:.3f} ms", flush=True)
