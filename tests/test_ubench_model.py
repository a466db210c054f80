"""scripts/ubench_tcgen05.cu is the hardware experiment that settles how tcgen05 shared-memory descriptors address operands the
kernels write themselves.  Its host-model build (-DUBENCH_HOST_MODEL) stages the operands and builds the descriptors with the
same helpers as the CUDA kernel and walks them according to the reading under test: every "as read" variant must reproduce the
product and every "exchanged" variant must not.  This keeps the experiment itself honest -- a FAIL on the GPU then means the
reading is wrong, not the staging code."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ubench_host_model(tmp_path):
    exe = str(tmp_path / 'ubench_model')
    src = os.path.join(ROOT, 'scripts', 'ubench_tcgen05.cu')
    r = subprocess.run(['g++', '-x', 'c++', '-O2', '-DUBENCH_HOST_MODEL', '-o', exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    lines = [l for l in r.stdout.splitlines() if ': model ' in l]
    assert len(lines) == 24, r.stdout
    assert r.returncode == 0 and 'UNEXPECTED' not in r.stdout, r.stdout
    assert sum('as read' in l and 'model PASS' in l for l in lines) == 12
    assert sum('exchanged' in l and 'model FAIL' in l for l in lines) == 12
