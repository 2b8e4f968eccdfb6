"""Imports the read-only Python reference (romeric/florence) for fixture generation ONLY.

Used by tests/golden/make_golden.py in the build container; never at test/bench time on the GPU box.
The reference's Fastor-free Cython modules are built from a scratch copy (recipe below); the seven
extension modules that need the absent Fastor headers / cblas.h are replaced by stubs that raise if
called, so every number written to tests/golden/ comes from the reference's own code:
its pure-numpy formulas and its Fastor-free native code.
"""
import os
import shutil
import subprocess
import sys
import types
import warnings

REF_SRC = "/root/reference"
REF_COPY = "/tmp/flref"

_BUILD_SCRIPT = r'''
import os, numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
mods = [
 ("Florence/Tensor/Numeric.pyx", "Florence.Tensor.Numeric"),
 ("Florence/Tensor/LinAlg.pyx", "Florence.Tensor.LinAlg"),
 ("Florence/FunctionSpace/JacobiPolynomials/JacobiPolynomials.pyx", "Florence.FunctionSpace.JacobiPolynomials.JacobiPolynomials"),
 ("Florence/FunctionSpace/OneDimensional/_OneD/LineBP.pyx", "Florence.FunctionSpace.OneDimensional.LineBP"),
 ("Florence/FiniteElements/Assembly/_Assembly_/ComputeSparsityPattern.pyx", "Florence.FiniteElements.Assembly.ComputeSparsityPattern"),
 ("Florence/FiniteElements/Assembly/_Assembly_/SparseAssemblyNative.pyx", "Florence.FiniteElements.Assembly.SparseAssemblyNative"),
 ("Florence/FiniteElements/Assembly/_Assembly_/RHSAssemblyNative.pyx", "Florence.FiniteElements.Assembly.RHSAssemblyNative"),
 ("Florence/VariationalPrinciple/_GeometricStiffness_/_GeometricStiffness_.pyx", "Florence.VariationalPrinciple._GeometricStiffness_"),
 ("Florence/VariationalPrinciple/_ConstitutiveStiffness_/DisplacementApproachIndices.pyx", "Florence.VariationalPrinciple.DisplacementApproachIndices"),
 ("Florence/VariationalPrinciple/_ConstitutiveStiffness_/DisplacementPotentialApproachIndices.pyx", "Florence.VariationalPrinciple.DisplacementPotentialApproachIndices"),
 ("Florence/MeshGeneration/HigherOrderMeshing/NPFROMFILE_Loop_Cyhton.pyx", "Florence.MeshGeneration.HigherOrderMeshing.NPFROMFILE_Loop_Cyhton"),
]
exts = [Extension(name, [m], include_dirs=[numpy.get_include(), os.path.dirname(m), "Florence/Tensor",
        "Florence/FunctionSpace/JacobiPolynomials"], language="c++", extra_compile_args=["-O2", "-std=c++14", "-w"]) for m, name in mods]
setup(ext_modules=cythonize(exts, language_level=3, force=True), script_args=["build_ext", "--inplace", "-j", "8"])
'''

STUBBED = [
    "Florence.FiniteElements.LocalAssembly._KinematicMeasures_",
    "Florence.VariationalPrinciple._MassIntegrand_",
    "Florence.VariationalPrinciple._ConstitutiveStiffnessDF_",
    "Florence.VariationalPrinciple._TractionDF_",
    "Florence.VariationalPrinciple._ConstitutiveStiffnessDPF_",
    "Florence.VariationalPrinciple._TractionDPF_",
    "Florence.VariationalPrinciple._ConstitutiveStiffnessLaplacian_",
]


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name in ("__path__", "__file__", "__spec__", "__loader__", "__all__"):
            raise AttributeError(name)
        modname = self.__name__

        def _raise(*a, **k):
            raise NotImplementedError("stub of Fastor/cblas-dependent reference module %s.%s" % (modname, name))
        return _raise


def load():
    if not os.path.isdir(os.path.join(REF_COPY, "Florence")):
        os.makedirs(REF_COPY, exist_ok=True)
        shutil.copytree(os.path.join(REF_SRC, "Florence"), os.path.join(REF_COPY, "Florence"))
    marker = os.path.join(REF_COPY, ".built")
    if not os.path.exists(marker):
        with open(os.path.join(REF_COPY, "build_ext.py"), "w") as f:
            f.write(_BUILD_SCRIPT)
        subprocess.check_call([sys.executable, "build_ext.py"], cwd=REF_COPY, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        open(marker, "w").write("ok")
    if REF_COPY not in sys.path:
        sys.path.insert(0, REF_COPY)
    for name in STUBBED:
        sys.modules.setdefault(name, _Stub(name))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import Florence
    return Florence
