"""bindings/node/: the N-API addon cannot be loaded here (no Node), but it must at least be valid C against the N-API
declarations it uses and the C ABI header, and the TypeScript wrapper must only call what the addon exports."""
import os
import re
import subprocess

from conftest import ROOT

NODE = os.path.join(ROOT, "bindings", "node")


def test_addon_compiles_against_the_napi_declarations_and_the_abi_header():
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "stubs"),
           os.path.join(NODE, "addon.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_addon_validates_every_typed_array_before_the_abi_call():
    src = open(os.path.join(NODE, "addon.c")).read()
    # no raw, unchecked typed-array accessor is left: every pointer comes from typed_arg()
    assert "typed_data(" not in src
    body = src[src.index("static napi_value js_create("):]
    assert body.count("napi_get_typedarray_info") == 0
    for fn in ("yalps_solve_batch_basis", "yalps_solve(", "yalps_multi_solve_many"):
        assert fn in src


def test_typescript_wrapper_calls_only_exported_addon_functions():
    ts = open(os.path.join(NODE, "yalps_b200.ts")).read()
    addon = open(os.path.join(NODE, "addon.c")).read()
    exported = set(re.findall(r'\{"(\w+)", NULL, js_', addon))
    used = set(re.findall(r"\bnative\.(\w+)\(", ts))
    assert used and used <= exported, (used, exported)
    # solveMany is implemented (round 1 shipped a stub that never called the addon)
    assert "native.solveMany(" in ts and "native.simplex(" in ts
