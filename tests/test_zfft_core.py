"""The float64 FFT arithmetic of the pruned z-pass kernel (bskit_b200/csrc/zfft_core.h) is plain
C++: compile the CPU harness with g++ and check pack + multi-radix Stockham stages against a
naive inverse DFT for every supported length, cropped and full-spectrum inputs."""
import os
import subprocess

from conftest import ROOT


def test_zfft_core_against_naive_dft(tmp_path):
    exe = str(tmp_path / "zfft_harness")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "zfft_harness.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    worst = float(out.stdout.strip().splitlines()[-1].split()[1])
    assert worst < 1e-12
    assert out.stdout.count("rel_err") == 12
