"""Time kernel variants back to back on one GPU box: ab_bench.py A B ... (variants/NAME.so, see build_variant.py).
Each variant runs bench.py in its own process, the whole list twice (drift shows up as a difference between rounds)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

if __name__ == "__main__":
    names = sys.argv[1:]
    for rnd in range(2):
        for n in names:
            env = dict(os.environ)
            if n != "product":
                env["LEVINS_LIB"] = os.path.join(HERE, "variants", n + ".so")
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu",
                                "--no-stream", "--e2e-steps", "2"], env=env, capture_output=True, text=True)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                ro = d["roofline"]
                print(f"{n:10s} round {rnd}: step {d['ms_per_step']:.4f} ms  k_deskew_bin {1e3 * ro['launch_ms']:.1f} us  frac {ro['frac']:.3f}  "
                      f"cfg3 {d['also']['k_deskew_bin_hbm_frac'] if d.get('also') else None}  "
                      f"f32 {d['single_precision_mode']['k_deskew_bin_launch_ms'] * 1e3 if d.get('single_precision_mode') else None} us "
                      f"step {d['single_precision_mode']['ms_per_step'] if d.get('single_precision_mode') else None}  "
                      f"share {[round(v, 3) for v in ro['kernel_time_share'].values()]}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(n, "failed", ex, r.stdout[-500:], r.stderr[-1500:], flush=True)
