"""TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference extensions into oracle/_ref/.

Compiles the three hot-path torch extensions of mit-han-lab/bevfusion straight from
their sources under /root/reference (nothing is copied into this repo):

  bev_pool_ext_ref     <- mmdet3d/ops/bev_pool/src/{bev_pool_cpu.cpp,bev_pool_cuda.cu}
  voxel_layer_ref      <- mmdet3d/ops/voxel/src/{voxelization.cpp,voxelization_cpu.cpp,
                          voxelization_cuda.cu,scatter_points_cpu.cpp,scatter_points_cuda.cu}
  sparse_conv_ext_ref  <- mmdet3d/ops/spconv/src/*.{cc,cu} (+ include/)

The recipe is a short torch.utils.cpp_extension.load() call per module (the same thing
the reference's setup.py:8-48 does, minus its sm_70..86 arch list); we do NOT run the
reference's setup.py.  CUDA code is cross-compiled for sm_100 so the very same .so runs
the reference's GPU kernels on the B200 box (exact oracle + on-box GPU baseline) and the
reference's CPU paths anywhere.

Outputs go to oracle/_ref/ only (git-ignored, but shipped to the GPU box by gpurun).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may load these modules.
"""
import os
import sys
import shutil

REF = os.environ.get("BEVFUSION_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

MODULES = ("bev_pool_ext_ref", "voxel_layer_ref", "sparse_conv_ext_ref")


def built(name):
    return os.path.exists(os.path.join(OUT, name + ".so"))


def build(verbose=False, only=None):
    if not os.path.isdir(REF):
        return False  # GPU box: use the prebuilt files
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    half = ["-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
            "-D__CUDA_NO_HALF2_OPERATORS__"]

    def one(name, sources, **kw):
        if only and name not in only:
            return
        if built(name):
            return
        bdir = os.path.join("/tmp", "bevfusion_ref_build", name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name, sources=sources, build_directory=bdir, verbose=verbose,
             with_cuda=True, is_python_module=False, **kw)
        shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))

    B = REF + "/mmdet3d/ops/bev_pool/src/"
    one("bev_pool_ext_ref", [B + "bev_pool_cpu.cpp", B + "bev_pool_cuda.cu"],
        extra_cuda_cflags=half)
    S = REF + "/mmdet3d/ops/voxel/src/"
    one("voxel_layer_ref",
        [S + f for f in ["voxelization.cpp", "scatter_points_cpu.cpp", "scatter_points_cuda.cu",
                         "voxelization_cpu.cpp", "voxelization_cuda.cu"]],
        extra_cflags=["-DWITH_CUDA"], extra_cuda_cflags=["-DWITH_CUDA"] + half)
    R = REF + "/mmdet3d/ops/spconv/"
    one("sparse_conv_ext_ref",
        [R + "src/" + f for f in ["all.cc", "reordering_cpu.cc", "reordering_cuda.cu",
                                  "indice_cpu.cc", "indice_cuda.cu", "maxpool_cpu.cc",
                                  "maxpool_cuda.cu"]],
        extra_include_paths=[R + "include"], extra_cflags=["-w", "-std=c++17"],
        extra_cuda_cflags=["-w", "-std=c++17"])
    return True


def load_ref(name):
    """Import a prebuilt reference module from oracle/_ref (torch must be imported first)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols)
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (run python oracle/build_ref.py in the build container)")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("built" if ok else "reference tree absent; nothing built", [m for m in MODULES if built(m)])
