"""The Rust FFI crate (bindings/rust/sim-b200-sys) declares exactly what include/simb.h declares and what
libsimb200.so exports.  (No Rust toolchain in the image: the crate is generated from the header by
tools/gen_rust_sys.py and checked here as text.)"""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYS = os.path.join(ROOT, "bindings", "rust", "sim-b200-sys", "src", "lib.rs")


def _header_functions():
    header = open(os.path.join(ROOT, "include", "simb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return set(re.findall(r"\b(simb_[a-z0-9_]+)\s*\(", header))


def test_sys_crate_is_generated_from_the_header():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_sys_crate_declares_every_exported_function():
    rust = open(SYS).read()
    declared = set(re.findall(r"pub fn (simb_[a-z0-9_]+)\(", rust))
    want = _header_functions()
    assert declared == want, (sorted(want - declared), sorted(declared - want))
    lib = C.CDLL(os.path.join(ROOT, "sim_b200", "libsimb200.so"))
    assert not [n for n in sorted(declared) if not hasattr(lib, n)]
    # the multi-GPU statistics reduction of the north_star is reachable from Rust
    for name in ("simb_stat_partial", "simb_stat_partial_sqdev", "simb_steady_state", "simb_get_records",
                 "simb_get_trace_hash", "simb_synchronize"):
        assert name in declared


def test_sys_crate_struct_layouts():
    """Field lists of the #[repr(C)] structs against the ctypes mirrors the tests use (same order, same widths)."""
    from sim_b200.simulator import SimbConfig, _CI, _Counters
    rust = open(SYS).read()
    width = {"u8": 1, "u16": 2, "u32": 4, "i32": 4, "u64": 8, "f64": 8, "*const c_char": 8}

    def fields(name):
        body = re.search(r"pub struct %s \{(.*?)\}" % name, rust, flags=re.S).group(1)
        return re.findall(r"pub (\w+): ([^,]+),", body)

    for name, mirror in (("simb_config", SimbConfig), ("simb_counters", _Counters), ("simb_ci", _CI)):
        fs = fields(name)
        assert [f for f, _ in fs] == [f for f, _ in mirror._fields_], name
        assert [width[t] for _, t in fs] == [C.sizeof(t) for _, t in mirror._fields_], name
