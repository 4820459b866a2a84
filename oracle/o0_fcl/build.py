"""Builds oracle/_ref/o0_se3 (the reference's SE3RigidBodyScenario on real FCL) where the dependencies exist.

TEST INFRASTRUCTURE.  Needs: pkg-config entries for fcl, eigen3, assimp (ccd comes with fcl), the nigh headers
(NIGH_INCLUDE, or a `nigh` checkout next to the reference as its CMakeLists.txt:71 expects), a jilog.hpp (the reference
expects it on the include path; a two-line stand-in is generated when it is missing) and the reference tree
(MPLB_REFERENCE_ROOT, default /root/reference).  Prints the binary's path, or `unavailable: <reason>` and exits 3."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "_ref")
REF = os.environ.get("MPLB_REFERENCE_ROOT", "/root/reference")


def unavailable(why):
    print("unavailable: " + why)
    sys.exit(3)


def pkg(*args):
    return subprocess.run(["pkg-config"] + list(args), capture_output=True, text=True)


def main():
    if not shutil.which("pkg-config"):
        unavailable("pkg-config not found")
    for mod in ("fcl", "eigen3", "assimp"):
        if pkg("--exists", mod).returncode != 0:
            unavailable("pkg-config does not know %s" % mod)
    if not os.path.exists(os.path.join(REF, "include", "mpl", "demo", "se3_rigid_body_scenario.hpp")):
        unavailable("reference tree not found at %s" % REF)
    nigh = os.environ.get("NIGH_INCLUDE")
    for cand in ([nigh] if nigh else []) + [os.path.join(REF, "..", "nigh", "src"), os.path.join(REF, "nigh", "src")]:
        if cand and os.path.exists(os.path.join(cand, "nigh", "se3_space.hpp")):
            nigh = cand
            break
    else:
        unavailable("nigh headers not found (set NIGH_INCLUDE to the directory that holds nigh/se3_space.hpp)")
    os.makedirs(OUT_DIR, exist_ok=True)
    shim = os.path.join(OUT_DIR, "o0_include")
    os.makedirs(shim, exist_ok=True)
    have_jilog = any(os.path.exists(os.path.join(d, "jilog.hpp")) for d in (os.path.join(REF, "include"), nigh, "/usr/include", "/usr/local/include"))
    if not have_jilog:   # logging only: swallow the stream
        with open(os.path.join(shim, "jilog.hpp"), "w") as f:
            f.write("#pragma once\n#include <iostream>\nstruct JiNull { template <class T> JiNull &operator<<(const T &) { return *this; } };\n"
                    "#define JI_LOG(level) JiNull()\n")
    flags = pkg("--cflags", "fcl", "eigen3", "assimp").stdout.split()
    libs = pkg("--libs", "fcl", "assimp").stdout.split()
    exe = os.path.join(OUT_DIR, "o0_se3")
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(REF, "include"), "-I", nigh, "-I", shim] + flags + \
          [os.path.join(HERE, "o0_main.cpp"), "-o", exe] + libs + ["-lccd", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        unavailable("compile failed: " + r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "compile failed")
    print(exe)


if __name__ == "__main__":
    main()
