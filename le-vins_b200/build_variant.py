"""Build a kernel variant next to the product library: build_variant.py NAME [-DFLAG ...] -> variants/NAME.so (git-ignored).
Used with LEVINS_LIB=... to time variants back to back on the same GPU box."""
import os
import subprocess
import sys

import build as B

if __name__ == "__main__":
    name, defs = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(B.HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name + ".so")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run([B.nvcc_path()] + B.NVCC_FLAGS + defs + ["-o", out] + B.SRC, env=env, check=True)
    print(out)
