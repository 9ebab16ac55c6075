"""Stage the reference's own Python sources under oracle/_ref/ so that its numba kernels can be TIMED on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box; ``oracle/_ref/`` is listed in
``.gitignore`` (nothing of the reference enters the history) but not in ``.gpurunignore``, so -- like the built
``.so`` files -- it travels with the tree.  ``__graft_entry__.build()`` calls :func:`stage` whenever
``/root/reference`` is present; the files are byte-for-byte copies (``oracle/ref_shims.py`` supplies the few import
shims at run time, in memory).  Only ``bench.py --impl reference`` (and its ``cpu_baseline`` leg) use the staged tree,
through ``oracle/ref_numba.py``; without it they fall back to the C port and say ``kind: "port"``.

usage: python oracle/stage_ref.py [reference_root]"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
# the modules ref_shims.load() imports, and what they import from their own package
NEEDED = (
    "sampler/cpl_functions.py", "sampler/cpl_constant_functions.py", "sampler/temporal_functions.py",
    "sampler/background.py", "sampler/source.py", "sampler/source_function.py", "sampler/cpl_source.py",
    "sampler/constant_cpl.py", "sampler/sampler.py", "response/response.py", "lightcurve/lightcurve.py",
    "lightcurve/light_curve_storage.py", "instruments/gbm/gbm_lightcurve.py", "utils/numba_array.py",
    "utils/logging.py", "utils/meta.py", "utils/package_utils.py", "utils/interpolation.py", "utils/time_interval.py", "utils/sys_tools.py",
)


def stage(reference_root="/root/reference", dest=DEST):
    """copy the needed files (those that exist) of <reference_root>/cosmogrb into <dest>/cosmogrb; returns the count"""
    src_pkg = os.path.join(reference_root, "cosmogrb")
    if not os.path.isdir(src_pkg):
        return 0
    n = 0
    for rel in NEEDED:
        s = os.path.join(src_pkg, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(dest, "cosmogrb", rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or os.path.getmtime(s) > os.path.getmtime(d) or os.path.getsize(s) != os.path.getsize(d):
            shutil.copyfile(s, d)
        n += 1
    return n


def staged_root(dest=DEST):
    """the directory to hand to ref_shims (COSMOGRB_REFERENCE) if the staged tree is usable, else None"""
    return dest if os.path.exists(os.path.join(dest, "cosmogrb", "sampler", "cpl_functions.py")) else None


if __name__ == "__main__":
    print(stage(*(sys.argv[1:2])), "files staged under", DEST)
