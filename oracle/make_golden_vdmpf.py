#!/usr/bin/env python
"""Golden outputs of the reference's multi-point scheme and its plugins: compiles tests/cpp/vdmpf_parity.cpp (fss::Vdmpf Gen /
BatchEval: key digests, output shares, proofs) and tests/cpp/prp_cuckoo_parity.cpp (fss::prp::Aes128Feistel, fss::cuckoo_hash)
against the UNMODIFIED reference headers with g++ (CPU, OpenSSL; -DNDEBUG as the reference's release build) and writes what
they print to tests/golden/vdmpf_v1.txt and tests/golden/prp_cuckoo_v1.txt.  Needs /root/reference; run in the build container
only.  Test infrastructure."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("FSS_REFERENCE", "/root/reference")
GOMP = ["-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"]


def main():
    os.makedirs(os.path.join(HERE, "_ref"), exist_ok=True)
    for src, golden in (("vdmpf_parity.cpp", "vdmpf_v1.txt"), ("prp_cuckoo_parity.cpp", "prp_cuckoo_v1.txt")):
        exe = os.path.join(HERE, "_ref", src.replace(".cpp", "_ref"))
        subprocess.run(["g++", "-std=c++20", "-O1", "-fopenmp", "-w", "-DNDEBUG", "-x", "c++", "-I/usr/local/cuda/include",
                        "-I", os.path.join(REF, "include"), os.path.join(ROOT, "tests", "cpp", src), "-o", exe, "-lcrypto", *GOMP],
                       check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
        if "verify=0" in out or "bijection=0" in out:
            sys.exit(f"{src}: the reference does not verify its own run: the test program is broken")
        path = os.path.join(ROOT, "tests", "golden", golden)
        with open(path, "w") as f:
            f.write(out)
        print(f"{path}: {len(out.splitlines())} lines")


if __name__ == "__main__":
    main()
