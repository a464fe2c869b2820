"""Packaging (ref ``setup.py:14-28``: package list + the ``tmlauncher`` script).

Unlike the reference, the package carries native code: ``python setup.py build_ext --inplace`` (or
``python -m theanompi_b200.csrc.build``) compiles ``theanompi_b200/csrc/*.cu|cpp`` for sm_100a with nvcc into
``theanompi_b200/_tmpi_native.so`` — in-tree, one torch-free shared object bound with pybind11.
"""
from setuptools import Command, find_packages, setup
from setuptools.command.build_py import build_py


class BuildNative(Command):
    description = "compile the sm_100a kernels + runtime into theanompi_b200/_tmpi_native.so"
    user_options = [("inplace", "i", "ignored: the extension is always built in-tree"), ("force", "f", "rebuild even when up to date")]
    boolean_options = ["inplace", "force"]

    def initialize_options(self):
        self.inplace = True
        self.force = False

    def finalize_options(self):
        pass

    def run(self):
        from theanompi_b200.csrc import build as native_build
        native_build.build(force=bool(self.force))


class BuildPyWithNative(build_py):
    def run(self):
        try:
            self.run_command("build_ext")
        except Exception as e:  # noqa: BLE001 — nvcc missing: ship the Python layer, ops fail loudly on a GPU box
            print("warning: native extension not built (%r)" % (e,))
        super().run()


setup(
    name="theanompi_b200",
    version="0.1.0",
    description="B200-native data-parallel training framework (BSP / EASGD / GOSGD) with the Theano-MPI user surface",
    packages=find_packages(include=["theanompi_b200", "theanompi_b200.*"]),
    package_data={"theanompi_b200": ["_tmpi_native.so", "csrc/*.cu", "csrc/*.cuh", "csrc/*.cpp", "csrc/*.h", "bin/tmlauncher"]},
    scripts=["theanompi_b200/bin/tmlauncher"],
    cmdclass={"build_ext": BuildNative, "build_py": BuildPyWithNative},
    python_requires=">=3.9",
)
