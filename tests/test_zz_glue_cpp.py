"""The C++ OneFlow glue (of-spmm_b200/oneflow_glue/spmm_{op,kernels,functor,grad}.cpp) compiled against
the minimal framework stand-in in tests/mock_oneflow and driven like the reference's framework
drives a user op (tests/cpp/glue_harness.cpp): inference, SBP, REGISTER_USER_KERNEL predicates,
InferTmpSize on the CPU; OpKernel::Compute on a CUDA stream on the GPU box."""
import os
import subprocess

import pytest

import ofspmm_b200 as ofs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "of-spmm_b200", "oneflow_glue")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    exe = str(tmp_path / "glue_harness")
    libdir = os.path.dirname(ofs.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-DWITH_CUDA", "-I", os.path.join(ROOT, "tests", "mock_oneflow"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "cpp", "glue_harness.cpp"), os.path.join(GLUE, "spmm_op.cpp"),
           os.path.join(GLUE, "spmm_kernels.cpp"), os.path.join(GLUE, "spmm_functor.cpp"),
           os.path.join(GLUE, "spmm_grad.cpp"), "-o", exe, "-L", libdir, "-lofspmm_b200", f"-Wl,-rpath,{libdir}",
           "-L", os.path.join(CUDA, "lib64"), "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_glue_compiles_and_host_side_behaves(tmp_path):
    out = subprocess.run([_build(tmp_path), "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "glue host checks ok: 18 kernel registrations" in out.stdout
    assert "glue autograd checks ok" in out.stdout          # functors + OpExprGradFunction


@pytest.mark.gpu
def test_glue_kernels_run_through_opkernel_compute(tmp_path):
    out = subprocess.run([_build(tmp_path), "gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "glue gpu checks ok" in out.stdout
