"""CPU: static checks of the repository's own rules (no compute).

  * oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import
    it -- in particular nothing under fp8_quantization_b200/ or tools/ does;
  * the product package has no CPU fallback: no module of it imports numpy-based or torch-eager restatements of the
    quantiser, and every op wrapper goes through ``_lib.lib()``;
  * nothing that runs on the GPU box (the -m gpu tests, smoke(), bench.py) reads /root/reference;
  * the C header documents, for every entry point, the reference function it replaces (file:line).
"""
import ast
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(*rel_dirs):
    for rel in rel_dirs:
        base = os.path.join(ROOT, rel)
        for dirpath, _, names in os.walk(base):
            if "__pycache__" in dirpath:
                continue
            for n in names:
                if n.endswith(".py"):
                    yield os.path.join(dirpath, n)


def _imports(path):
    """Top-level module names imported anywhere in the file (function bodies included)."""
    tree = ast.parse(open(path).read(), filename=path)
    mods = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            mods.update(a.name.split(".")[0] for a in node.names)
        elif isinstance(node, ast.ImportFrom) and node.level == 0 and node.module:
            mods.add(node.module.split(".")[0])
    return mods


def _path_literals(text, needle):
    """String constants (docstrings excluded) of a source text that contain ``needle``."""
    tree = ast.parse(text)
    doc = set()
    for node in ast.walk(tree):
        if isinstance(node, (ast.Module, ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)) and node.body:
            first = node.body[0]
            if isinstance(first, ast.Expr) and isinstance(first.value, ast.Constant) and isinstance(first.value.value, str):
                doc.add(id(first.value))
    return [n.value for n in ast.walk(tree) if isinstance(n, ast.Constant) and isinstance(n.value, str)
            and needle in n.value and id(n) not in doc]


def test_only_tests_smoke_and_bench_import_the_oracle():
    offenders = [p for p in _py_files("fp8_quantization_b200", "tools") if "oracle" in _imports(p)]
    assert not offenders, f"oracle/ imported outside tests/, smoke() and bench.py: {offenders}"
    # the product package does not even mention the oracle's modules or build outputs
    for p in _py_files("fp8_quantization_b200"):
        text = open(p).read()
        assert "fp8_oracle" not in text and "oracle/_build" not in text and "host_emul" not in text, p
        assert "fp8fq_sim" not in text and "host_sim" not in text, p   # the host simulation is test infrastructure


def test_bench_and_entry_touch_the_oracle_only_where_allowed():
    """bench.py: only inside the CPU-baseline / reference-arm function; __graft_entry__: only inside smoke()
    (build() compiles the checker, which is not using it)."""
    for fname, allowed in (("bench.py", None), ("__graft_entry__.py", {"smoke"})):
        tree = ast.parse(open(os.path.join(ROOT, fname)).read())
        for node in tree.body:  # module level: no oracle import
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                names = [a.name for a in node.names] + [getattr(node, "module", "") or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), f"{fname}: module-level oracle import"
        holders = set()
        for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
            for node in ast.walk(fn):
                if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                    holders.add(fn.name)
                if isinstance(node, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in node.names):
                    holders.add(fn.name)
        if allowed is not None:
            assert holders <= allowed, f"{fname}: oracle imported in {holders - allowed}"
        else:
            assert holders, "bench.py must time the oracle as its cpu_baseline / reference arm"
            assert all("cpu" in h or "reference" in h or "baseline" in h for h in holders), holders


def test_gpu_side_code_never_reads_the_reference_checkout():
    """/root/reference does not exist on the GPU box.  The only places that may name it are the loader that is
    guarded by reference_available() and the golden-vector generator (both run in the build container only)."""
    allowed = {os.path.join("oracle", "reference_loader.py"), os.path.join("tests", "golden", "make_golden.py"),
               os.path.join("tests", "test_repo_rules.py")}
    offenders = []
    for p in list(_py_files("fp8_quantization_b200", "tests", "tools", "oracle")) + [
            os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        rel = os.path.relpath(p, ROOT)
        if rel in allowed:
            continue
        text = open(p).read()
        if _path_literals(text, "/root/reference"):
            # a use is fine when it is guarded: the file must consult reference_available() / os.path.exists
            if "reference_available" not in text and "os.path.exists(\"/root/reference\")" not in text:
                offenders.append(rel)
    assert not offenders, f"unguarded use of /root/reference: {offenders}"


def test_every_op_wrapper_calls_the_library():
    """ops.py has no arithmetic of its own: every public function that returns quantised data reaches lib()."""
    src = open(os.path.join(ROOT, "fp8_quantization_b200", "ops.py")).read()
    tree = ast.parse(src)
    hot = {"prepare", "set_range_prepare", "fake_quant", "fake_quant_codes", "bn_fold", "bn_pack", "bn_act_quant",
           "bn_quant_add_act_quant", "fake_quant_multi", "add_act_quant", "fake_quant_backward", "uniform_prepare",
           "uniform_quant", "space_to_depth2", "max_pool2d_channels_last", "minmax", "estimate_prepare",
           "bn_act_estimate_prepare", "mse_grid", "fake_quant_host"}
    seen = set()
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in hot:
            seen.add(fn.name)
            calls = [n for n in ast.walk(fn) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute)
                     and isinstance(n.func.value, ast.Call) and getattr(n.func.value.func, "id", "") == "lib"]
            assert calls, f"ops.{fn.name} does not call libfp8fq.so"
            assert all(c.func.attr.startswith("fp8fq_") for c in calls)
    assert seen == hot, hot - seen


def test_header_cites_the_reference_for_every_entry_point():
    text = open(os.path.join(ROOT, "include", "fp8fq.h")).read()
    # split into (comment, declaration) pairs: each exported function must sit under a comment that either cites a
    # reference file:line or says that it has no counterpart / is introspection
    decls = re.findall(r"/\*(.*?)\*/\s*((?:(?:int64_t|int|const char\*)\s+fp8fq_[a-z0-9_]+\s*\([^;]*\);\s*)+)", text, flags=re.S)
    covered = set()
    for comment, block in decls:
        names = re.findall(r"(fp8fq_[a-z0-9_]+)\s*\(", block)
        cites = re.search(r"[a-z_/0-9]+\.py:\d+", comment) is not None
        meta = re.search(r"introspection|Host helpers|no counterpart|Number of kernel launches|End-to-end entry point",
                         comment) is not None
        assert cites or meta, f"{names}: the comment above cites no reference file:line"
        covered.update(names)
    all_syms = set(re.findall(r"\b(fp8fq_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    missing = all_syms - covered
    # typedef'd struct consumers are declared right after the struct, under the same comment
    assert missing <= {"fp8fq_fake_quant_multi_f32"}, missing
