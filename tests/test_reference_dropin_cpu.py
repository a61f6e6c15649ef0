"""The reference's own scripts against this repo's module shims (north star: "drops into py/flash_helpers/test and
tools/benchmark unchanged").  CPU only; skipped where /root/reference does not exist (the GPU box).

* every name the reference scripts import from `flash_attention` / `flash_helpers.*` exists here
  (/root/reference/tools/benchmark/pt_bench.py:18-33, run_kernels.py:6-20, ncu_bench.py:11-19,
  py/flash_helpers/test/test.py:3-14, tools/debug/sanity_check.py);
* the scripts that have no CUDA work at module level import for real, with the repo (and, for the two third-party
  packages this image lacks, flash_helpers/_compat) on PYTHONPATH.
"""
import ast
import importlib
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
SCRIPTS = ["tools/benchmark/pt_bench.py", "tools/benchmark/run_kernels.py", "tools/benchmark/ncu_bench.py",
           "py/flash_helpers/test/test.py", "tools/debug/sanity_check.py"]

pytestmark = pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the build container")


def imported_names(path):
    tree = ast.parse(path.read_text())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] in ("flash_helpers",
                                                                                              "flash_attention"):
            out += [(node.module, a.name) for a in node.names]
        elif isinstance(node, ast.Import):
            out += [(a.name, None) for a in node.names if a.name.split(".")[0] in ("flash_helpers", "flash_attention")]
    return out


@pytest.mark.parametrize("script", SCRIPTS)
def test_every_imported_name_exists(script):
    names = imported_names(REF / script)
    assert names, script
    for module, name in names:
        mod = importlib.import_module(module)
        if name is not None:
            assert hasattr(mod, name), f"{script} imports {name} from {module}"


def test_operator_attributes_used_by_the_scripts():
    import flash_attention

    assert callable(flash_attention.forward) and callable(flash_attention.forward_timed)


@pytest.mark.parametrize("script", ["tools/benchmark/run_kernels.py", "tools/benchmark/ncu_bench.py",
                                    "py/flash_helpers/test/test.py"])
def test_script_imports_unchanged(script):
    """Import the unmodified file as a module (its `main()` / unittest.main() sits behind __name__ == '__main__').
    pt_bench.py and sanity_check.py allocate CUDA memory at import time and are covered by the name check only."""
    extra = []
    for pkg in ("prettytable", "parameterized"):
        try:
            importlib.import_module(pkg)
        except ImportError:
            extra = [str(ROOT / "flash_helpers" / "_compat")]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT)] + extra))
    code = ("import importlib.util, sys; spec = importlib.util.spec_from_file_location('ref_script', sys.argv[1]); "
            "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); print('imported')")
    p = subprocess.run([sys.executable, "-c", code, str(REF / script)], capture_output=True, text=True, env=env,
                       timeout=300, cwd="/tmp")
    assert p.returncode == 0 and "imported" in p.stdout, p.stderr[-2000:]


def test_parameterized_shim_expands_like_the_package():
    sys.path.insert(0, str(ROOT / "flash_helpers" / "_compat"))
    try:
        import unittest

        mod = importlib.import_module("parameterized")

        class T(unittest.TestCase):
            @mod.parameterized.expand([("a b", 1), ("c", 2)], skip_on_empty=True)
            def test_x(self, name, v):
                assert v in (1, 2)

        names = unittest.TestLoader().getTestCaseNames(T)
        assert len(names) == 2 and all(n.startswith("test_x_") for n in names)
        assert unittest.TextTestRunner(stream=open(os.devnull, "w")).run(
            unittest.TestLoader().loadTestsFromTestCase(T)).wasSuccessful()
    finally:
        sys.path.pop(0)
        sys.modules.pop("parameterized", None)
