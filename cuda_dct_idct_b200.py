"""Import shim: the package directory is named ``cuda-dct-idct_b200`` (not a valid Python
identifier), so ``import cuda_dct_idct_b200`` loads it from there under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda-dct-idct_b200")
_spec = importlib.util.spec_from_file_location(
    "cuda_dct_idct_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cuda_dct_idct_b200"] = _mod
_spec.loader.exec_module(_mod)

if __name__ == "__main__":  # `python -m cuda_dct_idct_b200 in out [k]`: runpy resolves the name to this file
    from cuda_dct_idct_b200.__main__ import main

    sys.exit(main(sys.argv))
