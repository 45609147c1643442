"""Time kernel variants back to back on one GPU box: ab_bench.py A B ... (variants/NAME.so, see build_variant.py).
Each variant runs bench.py in its own process, the whole list twice (drift shows up as a difference between rounds)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")]
    extra = [a for a in sys.argv[1:] if a.startswith("--")]  # e.g. --wave=8
    os.environ.setdefault("LEVINS_BATCH_CACHE", "/tmp/levins_batch_cache")
    for rnd in range(2):
        for n in names:
            env = dict(os.environ)
            lib, _, flags = n.partition(":")  # NAME or NAME:--wave=8,--keyframes=16
            if lib != "product":
                env["LEVINS_LIB"] = os.path.join(HERE, "variants", lib + ".so")
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu",
                                "--no-stream", "--no-cfg5", "--e2e-steps", "2"] + extra + [f for f in flags.split(",") if f],
                               env=env, capture_output=True, text=True)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                ro = d["roofline"]
                al = d.get("also") or {}
                print(f"{n:24s} round {rnd}: step {d['ms_per_step']:.4f} ms  k_deskew_bin {1e3 * ro['launch_ms']:.1f} us  frac {ro['frac']:.3f}  "
                      f"whole {ro['whole_step_frac']:.3f}  share {[round(v, 3) for v in ro['kernel_time_share'].values()]}  "
                      f"cfg2: step {al.get('ms_per_step')} frac {al.get('k_deskew_bin_hbm_frac')} kernels {al.get('kernel_ms')}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(n, "failed", ex, r.stdout[-500:], r.stderr[-1500:], flush=True)
