"""CPU: the register FFT codelets of opencorr_b200/csrc/fft_codelet.cuh (used by fftcc2d_reg.cu / fftcc3d_reg.cu) compiled
for the host and checked against a double-precision DFT for every supported length."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_fft_codelets_against_dft(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fft_codelet_host_test")
    cmd = [nvcc, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "opencorr_b200", "csrc"), "-o", exe,
           os.path.join(ROOT, "tests", "native", "fft_codelet_host_test.cu")]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count("fwd rel err") == 16
