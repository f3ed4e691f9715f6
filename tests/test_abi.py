"""The C-ABI library loads and exports every symbol include/vsb200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import vector_store_b200 as v
    from importlib import import_module
    native = import_module("vector_store_b200.host.native")
    lib = ctypes.CDLL(v.lib_path())
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in vsb200.h but not exported"
    bound = sorted(n for n, _, _ in native.SYMBOLS)
    assert bound == declared, "ctypes binding and header drifted apart"


def test_version_and_no_cpu_fallback():
    import vector_store_b200 as v
    assert v.version().startswith("vsb200-")
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        # without a device the product must fail loudly, not fall back
        try:
            v.GpuIndex(8)
        except v.VsbError as e:
            assert e.status == 6  # VSB_ECUDA
        else:
            raise AssertionError("GpuIndex was created without a CUDA device")
        # the partition set (vsb_set_*) validates its options on a real handle: same loud failure
        try:
            v.IndexSet(8)
        except v.VsbError as e:
            assert e.status == 6
        else:
            raise AssertionError("IndexSet was created without a CUDA device")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vector-store_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace(
                    "oracle/exact.c", "").replace("oracle/graph_oracle.py", ""), f"{f} references oracle/"


def _build_c_client(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_client")
    libdir = os.path.join(ROOT, "vector-store_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_client.c"), "-o", exe, "-L", libdir, "-lvsb200", "-lm",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_header_compiles_as_c_and_links(tmp_path):
    """include/vsb200.h is valid C99 and a pure-C program links against libvsb200.so."""
    import subprocess
    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert r.returncode == 0, r.stdout + r.stderr
        assert "scenario passed" in r.stdout
    else:
        assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)  # loud failure, no CPU fallback


def test_struct_layouts_match_the_ctypes_mirror(tmp_path):
    """sizeof() of every struct in include/vsb200.h (compiled as C) equals the ctypes mirror's, and vsb_stats
    lists the same fields in the same order — the binding cannot drift from the header unnoticed."""
    import ctypes
    import re
    import subprocess
    from importlib import import_module
    native = import_module("vector_store_b200.host.native")
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "vsb200.h"\nint main(void){printf("%zu %zu %zu\\n", '
                   'sizeof(vsb_options), sizeof(vsb_search_params), sizeof(vsb_stats));return 0;}\n')
    exe = str(tmp_path / "sizes")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(native.VsbOptions), ctypes.sizeof(native.VsbSearchParams), ctypes.sizeof(native.VsbStats)]
    assert got == want, (got, want)
    header = open(os.path.join(ROOT, "include", "vsb200.h")).read()
    body = header[:header.index("} vsb_stats;")]
    body = body[body.rindex("typedef struct"):]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in re.findall(r"uint64_t\s+([^;]+);", body):
        names += [n.strip() for n in decl.split(",")]
    assert names == [n for n, _ in native.VsbStats._fields_], (names, [n for n, _ in native.VsbStats._fields_])


def _header_struct_fields(name):
    text = open(os.path.join(ROOT, "include", "vsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    body = next(chunk.split("}")[0] for chunk in text.split("typedef struct {")[1:]
                if re.match(r"\s*" + name + r"\s*;", chunk.split("}", 1)[1]))
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        fields += [re.sub(r"\[.*?\]", "", n).strip() for n in names.split(",")]
    return fields


def test_rust_ffi_declares_every_header_symbol():
    """integration/vsb200-sys/src/lib.rs (uncompiled here: no cargo) stays in step with the header: every entry point
    declared, every struct with the same field names in the same order."""
    rs = open(os.path.join(ROOT, "integration", "vsb200-sys", "src", "lib.rs")).read()
    declared_rs = sorted(set(re.findall(r"pub fn (vsb_[a-z0-9_]+)\s*\(", rs)))
    assert declared_rs == _header_symbols()
    for struct in ("vsb_options", "vsb_search_params", "vsb_stats", "vsb_build_stats"):
        body = re.search(r"pub struct " + struct + r" \{(.*?)\n\}", rs, flags=re.S).group(1)
        fields_rs = re.findall(r"pub ([a-z0-9_]+):", body)
        assert fields_rs == _header_struct_fields(struct), struct
    gpu = open(os.path.join(ROOT, "integration", "vs_index", "gpu.rs")).read()
    for call in re.findall(r"sys::(vsb_[a-z0-9_]+)\(", gpu):
        assert call in declared_rs, call
    for elision in ("(..)", ", ..,", "-> ...", "..., "):
        assert elision not in gpu, "gpu.rs must be complete source, no elisions"
