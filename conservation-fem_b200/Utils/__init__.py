"""Drop-in for the reference's ``Code/Utils`` package (``PYTHONPATH=.../Code`` ->
``PYTHONPATH=.../conservation-fem_b200``): same module, class and method names,
same argument meaning; the per-node Python loops run as CUDA kernels."""
