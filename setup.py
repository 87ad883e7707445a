"""Build entry point with the reference's own command line (reference setup.py:1-55 builds the CUDA extensions `rmsnorm`
and `swiglu_fused` with `python setup.py build_ext --inplace`).

Here the same command compiles the C-ABI library `llama-3.2-multimodal_b200/libl32ffn.so` (hand-written CUDA for sm_100a, no
torch headers, ~20 s) in-tree; the importable modules `rmsnorm` and `swiglu_fused` at the repository root bind it through
ctypes and keep the reference's module names and callables.
"""
import importlib.util
import os

from setuptools import Command, setup

ROOT = os.path.dirname(os.path.abspath(__file__))


def _load_build():
    spec = importlib.util.spec_from_file_location("l32_build", os.path.join(ROOT, "llama-3.2-multimodal_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class BuildNative(Command):
    description = "compile libl32ffn.so for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a)"
    user_options = [("inplace", "i", "accepted for compatibility with the reference's command line (always in-tree)"),
                    ("force", "f", "rebuild even if the sources did not change")]
    boolean_options = ["inplace", "force"]

    def initialize_options(self):
        self.inplace = False
        self.force = False

    def finalize_options(self):
        pass

    def run(self):
        print(_load_build().build(force=bool(self.force), verbose=False))


setup(
    name="llama32-b200-ffn",
    version="0.1.0",
    description="B200-native Add-RMSNorm + SwiGLU feed-forward hot path (drop-in for emmanuelalo52/LLaMA-3.2-Multimodal)",
    py_modules=["rmsnorm", "swiglu_fused"],
    packages=["llama32_b200"],
    cmdclass={"build_ext": BuildNative},
    python_requires=">=3.9",
)
