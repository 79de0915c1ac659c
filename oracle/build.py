"""Builds the test oracle (TEST INFRASTRUCTURE -- never imported by dexdeform_b200/).

* ``libmpm_oracle.so``  : gcc build of oracle/mpm_oracle.c (the CPU restatement)
* ``_ref/*.so``         : the unmodified reference, via oracle/build_ref.sh, when /root/reference is present
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def _stale(target, sources):
    return not os.path.isfile(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources if os.path.exists(s))


def build(force=False, verbose=True):
    src = os.path.join(HERE, "mpm_oracle.c")
    out = os.path.join(HERE, "libmpm_oracle.so")
    if force or _stale(out, [src]):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", src, "-o", out, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    ref_src = os.environ.get("REF", "/root/reference") + "/mpm/csrc/integrator.cu"
    ref_out = [os.path.join(HERE, "_ref", n) for n in ("libmaniskill_mpm.so", "libmaniskill_mpm_cpu.so")]
    if os.path.isfile(ref_src) and (force or any(_stale(o, [ref_src, os.path.join(HERE, "ref_cpu.cpp")]) for o in ref_out)):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return out


if __name__ == "__main__":
    build(force=True)
