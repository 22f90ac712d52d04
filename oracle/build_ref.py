"""TEST SCAFFOLDING -- builds the UNMODIFIED reference rasterizer into oracle/_ref/.

The reference hot path (diff-gaussian-rasterization, "DGR") is CUDA, so it cannot run
in the GPU-less build container, but it compiles here and the resulting .so travels to
the GPU box with the repo snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
It is used by tests/ and bench.py ONLY as (a) the GPU-side parity oracle that pins the
C restatement in oracle/gs_oracle.c and generates tests/golden/*.npz, and (b) the speed
baseline of `bench.py --impl reference`.  Nothing in manigaussian_b200/ imports it.

Sources are compiled WHERE THEY LIE under /root/reference -- nothing is copied.  Three
things the reference tree lacks are supplied from the command line instead of by edits:
  * glm (un-vendored submodule, DGR/.gitmodules:1-3)      -> -I oracle/glm_standin
  * <cstdint> for gcc>=13 in cuda_rasterizer/rasterizer_impl.h, and
  * the compile-time feature width NUM_CHANNELS_language_feature (config.h:16, stock 3)
    -> nvcc `-include oracle/ref_build_shim.h`: config.h's include guard is pre-defined
    and its four macros are defined by the shim, so the F=32 variant BASELINE.json's
    configs need is the same source with -DMGS_REF_FEATURE_CHANNELS=32.

Usage:  python oracle/build_ref.py [3 32 ...]     (default: 3 and 32)
Outputs: oracle/_ref/dgr_ref_f<F>/dgr_ref_f<F>.so  (pybind module exporting
rasterize_gaussians / rasterize_gaussians_backward / mark_visible, DGR/ext.cpp:14-18), and
oracle/_ref/python/diff_gaussian_rasterization/__init__.py -- the reference's own Python operator
(autograd.Function, GaussianRasterizer), INSTALLED there unmodified like `pip install` would (oracle/_ref is
build output: git-ignored, never committed), so that bench.py's reference arm drives the reference's own
wrapper on the GPU box where /root/reference does not exist (tests/util.py::load_reference_package).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get(
    "MGS_REFERENCE_DGR",
    "/root/reference/third_party/gaussian-splatting/submodules/diff-gaussian-rasterization")
OUT = os.path.join(HERE, "_ref")


def build_variant(F: int, verbose: bool = False):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    name = f"dgr_ref_f{F}"
    bdir = os.path.join(OUT, name)
    os.makedirs(bdir, exist_ok=True)
    srcs = [os.path.join(REF_ROOT, s) for s in (
        "cuda_rasterizer/rasterizer_impl.cu",
        "cuda_rasterizer/forward.cu",
        "cuda_rasterizer/backward.cu",
        "rasterize_points.cu",
        "ext.cpp")]
    for s in srcs:
        if not os.path.exists(s):
            raise FileNotFoundError(s)
    shim = os.path.join(HERE, "ref_build_shim.h")
    inc = ["-I" + os.path.join(HERE, "glm_standin")]
    mod = load(
        name=name,
        sources=srcs,
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "-lineinfo", "-include", shim, f"-DMGS_REF_FEATURE_CHANNELS={F}"] + inc,
        build_directory=bdir,
        verbose=verbose,
    )
    return mod


def main(argv):
    if not os.path.isdir(REF_ROOT):
        print(f"[build_ref] {REF_ROOT} not present (GPU box) -- using prebuilt oracle/_ref if any")
        return 0
    variants = [int(a) for a in argv] or [3, 32]
    import shutil
    pkg = os.path.join(OUT, "python", "diff_gaussian_rasterization")
    os.makedirs(pkg, exist_ok=True)
    shutil.copyfile(os.path.join(REF_ROOT, "diff_gaussian_rasterization", "__init__.py"), os.path.join(pkg, "__init__.py"))
    print(f"[build_ref] installed the reference's Python operator into {pkg}")
    for F in variants:
        m = build_variant(F, verbose=True)
        print(f"[build_ref] built {m.__file__}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
