"""Generates bindings/rust/src/sys.rs -- the raw `extern "C"` block a Rust maintainer of the reference would link
against libtakzero_b200.so -- from include/takzero_b200.h, so the two cannot drift apart
(tests/test_rust_bindings.py regenerates and compares).  The image has no rustc / cargo: the output is checked for
coverage and consistency only, never compiled.   python tools/gen_rust_bindings.py [--check]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "takzero_b200.h")
OUT = os.path.join(ROOT, "bindings", "rust", "src", "sys.rs")

SCALARS = {
    "int": "c_int", "float": "f32", "double": "f64", "char": "c_char", "void": "c_void", "size_t": "usize",
    "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "long long": "c_longlong",
    "tz_move_t": "tz_move_t", "tz_agent_fn": "tz_agent_fn", "tz_model_tensor_fn": "tz_model_tensor_fn",
}


def rust_type(c: str) -> str:
    c = " ".join(c.replace("*", " * ").split())
    stars = c.count("*")
    base = c.replace("*", "").strip()
    const = base.startswith("const ")
    base = base[6:].strip() if const else base
    base = base.replace("struct ", "")
    t = SCALARS.get(base, base)  # tz_* structs keep their names
    for i in range(stars):
        t = ("*const " if const and i == 0 else "*mut ") + t
    return t


def split_params(params: str):
    out = []
    for p in [x.strip() for x in params.split(",") if x.strip()]:
        if p == "void":
            continue
        m = re.match(r"(.+?)\s*(\w+)$", p)
        ctype, name = m.group(1), m.group(2)
        if name in ("fn", "move", "in", "type"):
            name += "_"
        out.append((name, rust_type(ctype)))
    return out


def parse(text: str):
    text_nc = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    funcs = []
    for m in re.finditer(r"TZ_API\s+([\w\s\*]+?)\s*\b(tz_\w+)\s*\(([^;]*?)\)\s*;", text_nc, flags=re.S):
        funcs.append((m.group(2), rust_type(m.group(1)), split_params(" ".join(m.group(3).split()))))
    structs = []
    for m in re.finditer(r"typedef struct (tz_\w+) \{(.*?)\} \1;", text_nc, flags=re.S):
        fields = []
        for line in m.group(2).split(";"):
            line = " ".join(line.split())
            if not line:
                continue
            fm = re.match(r"(.+?)\s*(\w+)((?:\[\w+\])*)$", line)
            ctype, name, dims = fm.group(1), fm.group(2), re.findall(r"\[(\w+)\]", fm.group(3))
            t = rust_type(ctype)
            for d in reversed(dims):
                t = f"[{t}; {d}]"
            fields.append((name, t))
        structs.append((m.group(1), fields))
    consts = []
    for m in re.finditer(r"#define (TZ_[A-Z_]+) (\d+)", text_nc):
        consts.append((m.group(1), m.group(2), "usize"))
    for m in re.finditer(r"enum\s*\{(.*?)\}", text_nc, flags=re.S):
        for item in m.group(1).split(","):
            item = item.strip()
            if "=" in item:
                k, v = [x.strip() for x in item.split("=")]
                consts.append((k, v, "c_int" if v.startswith("-") or k.startswith("TZ_E") or k.startswith("TZ_OK") else "u32"))
    callbacks = []
    for m in re.finditer(r"typedef void \(\*(tz_\w+)\)\((.*?)\);", text_nc, flags=re.S):
        callbacks.append((m.group(1), split_params(" ".join(m.group(2).split()))))
    return funcs, structs, consts, callbacks


def render() -> str:
    funcs, structs, consts, callbacks = parse(open(HEADER).read())
    o = ["//! Raw bindings of libtakzero_b200.so, GENERATED from include/takzero_b200.h by tools/gen_rust_bindings.py.",
         "//! UNVERIFIED: the image this was produced in has no Rust toolchain; the file is checked against the header",
         "//! (every exported tz_* function, struct and constant is present with the mapped types) but was never compiled.",
         "//! Link with `cargo:rustc-link-lib=dylib=takzero_b200` (see INTEGRATION.md).",
         "#![allow(non_camel_case_types, non_upper_case_globals, dead_code)]",
         "use core::ffi::{c_char, c_int, c_longlong, c_void};", "",
         "pub type tz_move_t = u16;", "#[repr(C)] pub struct tz_handle { _private: [u8; 0] }", ""]
    for name, val, ty in consts:
        o.append(f"pub const {name}: {ty} = {val};")
    o.append("")
    for name, fields in structs:
        o.append("#[repr(C)]\n#[derive(Clone, Copy)]\npub struct %s {" % name)
        for f, t in fields:
            o.append(f"    pub {f}: {t},")
        o.append("}")
    o.append("")
    for name, params in callbacks:
        args = ", ".join(f"{n}: {t}" for n, t in params)
        o.append(f"pub type {name} = Option<unsafe extern \"C\" fn({args})>;")
    o.append("")
    o.append('extern "C" {')
    for name, ret, params in funcs:
        args = ", ".join(f"{n}: {t}" for n, t in params)
        r = "" if ret == "c_void" else f" -> {ret}"
        o.append(f"    pub fn {name}({args}){r};")
    o.append("}")
    return "\n".join(o) + "\n"


if __name__ == "__main__":
    text = render()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(text)
    print(OUT)
