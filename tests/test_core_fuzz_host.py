"""Memory-safety fuzz of the host/device-shared transform cores (TEXT / UTF / EXE) under AddressSanitizer + UBSan: corrupted or random
inputs to the inverse walks and to the executable-header parsers must fail cleanly. The same functions run inside the GPU kernels,
where an out-of-bounds access would take the CUDA context down instead of returning an error."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "fuzz_cores.cpp")
EXE = os.path.join(ROOT, "tests", "host", "_build", "fuzz_cores")


def test_cores_survive_corrupted_input():
    if not os.path.exists(os.path.join(ROOT, "kanzi-go_b200", "csrc", "_gen", "kz_text_dict.inc")):
        pytest.skip("static dictionary not generated")
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", EXE, SRC],
                       cwd=os.path.dirname(SRC), capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("sanitizers not available: " + r.stderr[:200])
    assert r.returncode == 0, r.stderr
    r = subprocess.run([EXE, "400"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fuzz done" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
