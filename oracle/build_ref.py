"""TEST INFRASTRUCTURE ONLY -- build the REFERENCE's own CUDA op into oracle/_ref/.

Compiles nets/deform_conv/src/deform_conv_cuda{.cpp,_kernel.cu} from where they lie under
/root/reference (nothing is copied into the repo) with torch.utils.cpp_extension for sm_100a; the
only output is oracle/_ref/deform_conv_cuda*.so (git-ignored, shipped to the GPU box by gpurun).
It is used by tests/test_gpu_reference_op.py as a second, GPU-side checker and by
bench.py --compare-ref as the "kernel to beat"; never by the product.

The recipe does not run the reference's setup.py: it hands the two source files straight to
cpp_extension.load with our own flags.
"""
import glob
import os

REF = os.environ.get("AANET_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "deform_conv_cuda"


def built_path():
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return hits[0] if hits else None


def build_if_possible(verbose=False):
    src_dir = os.path.join(REF, "nets", "deform_conv", "src")
    srcs = [os.path.join(src_dir, "deform_conv_cuda.cpp"), os.path.join(src_dir, "deform_conv_cuda_kernel.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return built_path()
    hit = built_path()
    if hit and all(os.path.getmtime(hit) >= os.path.getmtime(s) for s in srcs):
        return hit
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils import cpp_extension
    cpp_extension.load(name=NAME, sources=srcs, build_directory=OUT, verbose=verbose,
                       extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-w"],
                       extra_cflags=["-O2", "-w"], is_python_module=False)
    return built_path()


def load():
    """Import the built module (GPU box or here); returns None when it was never built."""
    path = built_path()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_if_possible(verbose=True))
