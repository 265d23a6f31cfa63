"""The C ABI hosted from plain C (no Python, no PyTorch in the process): tests/c_abi/c_abi_host.c allocates with
cudaMalloc, runs one DMoL + KL + ELBO step through libblvm_b200.so and checks it against the C oracle in fp64."""
import os
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "c_abi", "c_abi_host.c")
OUT_DIR = os.path.join(ROOT, "tests", "c_abi", "_build")
LIB_DIR = os.path.join(ROOT, "benchmarking-lvms_b200", "lib")
ORACLE_DIR = os.path.join(ROOT, "oracle", "_build")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build():
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    exe = os.path.join(OUT_DIR, "c_abi_host")
    subprocess.run(["gcc", "-O2", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), SRC,
                    "-o", exe, os.path.join(LIB_DIR, "libblvm_b200.so"), os.path.join(ORACLE_DIR, "libblvm_oracle.so"),
                    "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{LIB_DIR}", f"-Wl,-rpath,{ORACLE_DIR}",
                    f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"], check=True)
    return exe


def test_c_host_compiles_against_the_header():
    """CPU: the header is valid C11 and the library resolves every symbol the C host uses (link step)."""
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_c_host_runs_one_step_against_the_oracle():
    res = subprocess.run([_build()], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "PASS" in res.stdout, res.stdout + res.stderr
