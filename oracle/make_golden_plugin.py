#!/usr/bin/env python
"""Golden output of the reference's class templates instantiated with the USER-DEFINED plugins of
tests/cpp/user_plugin.hpp (a two-lane integer group and a toy PRG): compiles oracle/plugin_ref_main.cpp against the
unmodified reference headers with g++ (CPU) and writes its output to tests/golden/plugin_user_v1.txt.  Needs
/root/reference; run in the build container only.  Test infrastructure."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("FSS_REFERENCE", "/root/reference")


def main():
    os.makedirs(os.path.join(HERE, "_ref"), exist_ok=True)
    exe = os.path.join(HERE, "_ref", "plugin_ref_main")
    subprocess.run(["g++", "-std=c++20", "-O1", "-fopenmp", "-x", "c++", "-I/usr/local/cuda/include", "-I", os.path.join(REF, "include"),
                    os.path.join(HERE, "plugin_ref_main.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    if "BAD" in out:
        sys.exit("the reference does not reconstruct with the user plugins: the plugins are broken")
    path = os.path.join(ROOT, "tests", "golden", "plugin_user_v1.txt")
    with open(path, "w") as f:
        f.write(out)
    print(f"{path}: {len(out.splitlines())} lines")


if __name__ == "__main__":
    main()
