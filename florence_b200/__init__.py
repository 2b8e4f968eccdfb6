"""florence_b200 -- B200 (sm_100a) back end for the element-assembly hot path of romeric/florence.

Layout: csrc/ (CUDA kernels + the C ABI of include/florence_b200.h), backend.py (device handle), assembly.py (the
reference-facing `_LowLevelAssembly_*` plug-in functions), time_integrator.py (device-resident explicit loop),
partition.py (element partitioning across GPUs), mesh.py (synthetic structured meshes and function-space tables).
"""
__version__ = "0.1.0"
