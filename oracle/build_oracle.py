#!/usr/bin/env python3
"""Build recipe for the CPU checker (TEST INFRASTRUCTURE, not product code).

Two things are built:

1. ``oracle/liblbm_oracle.so`` -- our own C restatement (oracle/lbm_oracle.c), always.

2. ``oracle/_ref/*`` -- the UNMODIFIED reference, compiled from where it lies under
   /root/reference (never copied into the repo; the derived variants are piped from
   ``sed`` straight into gcc's stdin).  Only possible where /root/reference exists
   (the build container); the built files travel to the GPU box with the snapshot
   (oracle/_ref/ is git-ignored, not gpurun-ignored).

   libref_f32_strict.so  d2q9-bgk.c, -O2 -ffp-contract=off, main renamed -> exports
                         timestep_new2, accelerate_flow, propagate, rebound,
                         collision, av_velocity ... for bit-exact unit comparisons
   libref_f32_fast.so    the reference's own Makefile flags (Makefile:6)
   libref_f32_omp.so     same flags + -fopenmp with the one-line annotation of the
                         row loop at d2q9-bgk.c:787: the
                         multi-core CPU baseline.  The reference has no pragma.
   libref_f64.so         mechanical float->double substitution (SURVEY.md section 7
                         step 1c): the generator of the golden files in check/
   d2q9-bgk_ref          the reference CLI binary, its own flags
   d2q9-bgk_ref_omp      CLI binary of the annotated copy
   d2q9-bgk_ref_f64      CLI binary of the fp64 substitution
   d2q9-bgk_ref_gpu      the reference with ONLY its step loop (d2q9-bgk.c:180-201) replaced
                         by the lbm_gpu_* binding of INTEGRATION.md (oracle/reference_binding.inc):
                         the reference's own initialise / calc_reynolds / write_values around
                         liblbm_b200.so -- the drop-in demonstration, tests/test_gpu_dropin.py
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("LBM_REFERENCE_DIR", "/root/reference")
REF_SRC = os.path.join(REF_DIR, "d2q9-bgk.c")
REF_MD5 = "93064ebaacec508bd2305e01ffd892a1"   # the line-787 annotation depends on this file
OUT_REF = os.path.join(HERE, "_ref")

# reference Makefile:6 is "-std=c99 -Wall -Ofast -mtune=native -march=native
# -funsafe-math-optimizations".  The built files travel to a GPU box whose host CPU may
# differ from the build container's (sapphirerapids here), so -march=native is replaced
# by the portable x86-64-v3 (AVX2+FMA) level; the hot loop is scalar code either way
# (0 vectorised loops, e000/hs000/vectorization.advisum).
REF_FLAGS = ["-std=c99", "-Wall", "-Ofast", "-mtune=generic", "-march=x86-64-v3",
             "-funsafe-math-optimizations"]
STRICT_FLAGS = ["-std=c99", "-O2", "-ffp-contract=off"]
SED_F64 = r's/\bfloat\b/double/g; s/sqrtf/sqrt/g; s/([0-9])\.f\b/\1.0/g; s/"%f\\n"/"%lf\\n"/g'
OMP_PRAGMA = "#pragma omp parallel for reduction(+:tot_u,tot_cells) private(y_s) schedule(static)"
SED_OMP = "787i " + OMP_PRAGMA


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode(errors="replace"))
        raise RuntimeError("command failed: %s" % (cmd,))
    return r


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle(force=False):
    """gcc the restatement into oracle/liblbm_oracle.so."""
    out = os.path.join(HERE, "liblbm_oracle.so")
    srcs = [os.path.join(HERE, "lbm_oracle.c"), os.path.join(HERE, "lbm_oracle_impl.h")]
    if not force and _newer(out, *srcs):
        return out
    _run(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
          "-Wall", srcs[0], "-lm", "-o", out])
    return out


def _gcc_from_sed(sed_args, gcc_args, out):
    """sed <REF_SRC> | gcc -x c - : no copy of the reference source is written."""
    if sed_args is None:
        _run(["gcc"] + gcc_args + [REF_SRC, "-lm", "-o", out])
        return
    sed = subprocess.Popen(["sed"] + sed_args + [REF_SRC], stdout=subprocess.PIPE)
    try:
        _run(["gcc"] + gcc_args + ["-x", "c", "-", "-lm", "-o", out], stdin=sed.stdout)
    finally:
        sed.stdout.close()
        sed.wait()


def reference_available():
    return os.path.isfile(REF_SRC)


def build_reference(force=False):
    """Compile the reference from /root/reference into oracle/_ref/ (if present)."""
    if not reference_available():
        return None
    with open(REF_SRC, "rb") as f:
        md5 = hashlib.md5(f.read()).hexdigest()
    if md5 != REF_MD5:
        raise RuntimeError("reference source changed (md5 %s): re-check the line-787 annotation" % md5)
    os.makedirs(OUT_REF, exist_ok=True)
    lib = ["-Dmain=ref_main", "-fPIC", "-shared"]
    targets = [
        ("libref_f32_strict.so", None, STRICT_FLAGS + lib),
        ("libref_f32_fast.so", None, REF_FLAGS + lib),
        ("libref_f32_omp.so", [SED_OMP], REF_FLAGS + ["-fopenmp"] + lib),
        ("libref_f64.so", ["-E", SED_F64], ["-std=c99", "-O2", "-ffp-contract=off"] + lib),
        ("d2q9-bgk_ref", None, REF_FLAGS),
        ("d2q9-bgk_ref_omp", [SED_OMP], REF_FLAGS + ["-fopenmp"]),
        ("d2q9-bgk_ref_f64", ["-E", SED_F64], ["-std=c99", "-O3", "-march=x86-64-v3"]),
    ]
    for name, sed_args, flags in targets:
        out = os.path.join(OUT_REF, name)
        if not force and _newer(out, REF_SRC, os.path.abspath(__file__)):
            continue
        _gcc_from_sed(sed_args, flags, out)
    build_dropin(force)
    return OUT_REF


def build_dropin(force=False):
    """The reference's main() with its step loop replaced by the C-ABI binding, linked
    against the product library (which must have been built already)."""
    root = os.path.dirname(HERE)
    pkg = os.path.join(root, "advanced-hpc-lbm_b200")
    lib = os.path.join(pkg, "liblbm_b200.so")
    inc = os.path.join(HERE, "reference_binding.inc")
    out = os.path.join(OUT_REF, "d2q9-bgk_ref_gpu")
    if not os.path.exists(lib):
        return None
    if not force and _newer(out, REF_SRC, inc, lib, os.path.abspath(__file__)):
        return out
    sed_args = ["-e", "180,201d", "-e", "179r " + inc, "-e", '57a #include "lbm_gpu.h"']
    gcc_args = ["-std=c99", "-O2", "-ffp-contract=off", "-I" + os.path.join(root, "include")]
    sed = subprocess.Popen(["sed"] + sed_args + [REF_SRC], stdout=subprocess.PIPE)
    try:
        _run(["gcc"] + gcc_args + ["-x", "c", "-", "-L" + pkg, "-llbm_b200",
                                   "-Wl,-rpath,$ORIGIN/../../advanced-hpc-lbm_b200", "-lm", "-o", out], stdin=sed.stdout)
    finally:
        sed.stdout.close()
        sed.wait()
    return out


def main():
    force = "--force" in sys.argv
    print("oracle:", build_oracle(force))
    ref = build_reference(force)
    print("reference:", ref if ref else "not available here (%s missing)" % REF_SRC)


if __name__ == "__main__":
    main()
