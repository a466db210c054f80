#!/usr/bin/env python
"""Stage the UNMODIFIED reference package next to the oracle so that it can be timed on the GPU box.  (test infrastructure)

    python oracle/stage_ref.py            # /root/reference/recbole_cdr -> oracle/_ref/recbole_cdr  (git-ignored)

The reference is pure Python; ``/root/reference`` exists only in the build container, while ``bench.py --impl reference`` and
the ``cpu_baseline`` leg run on the GPU box.  ``oracle/_ref/`` is listed in ``.gitignore`` (reference sources never enter this
repository's history) but not in ``.gpurunignore``, so the staged copy travels with the snapshot like the built ``.so``.
Only ``bench.py``'s CPU legs import it (over ``oracle/recbole_shim``, the stub of the un-vendored ``recbole==1.0.1``); nothing on
the product path does.  ``__graft_entry__.build()`` calls ``stage()`` whenever ``/root/reference`` is present.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('XDR_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')


def stage(verbose=False):
    """Copy the reference's python package verbatim; returns the destination or None when the reference is not mounted."""
    src = os.path.join(REF, 'recbole_cdr')
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DST, 'recbole_cdr')
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', 'dataset_example'))
    with open(os.path.join(DST, 'README'), 'w') as f:
        f.write('Verbatim copy of /root/reference/recbole_cdr made by oracle/stage_ref.py (git-ignored; used only by the CPU legs '
                'of bench.py).\n')
    if verbose:
        print('staged', dst)
    return dst


if __name__ == '__main__':
    print(stage(verbose=True))
