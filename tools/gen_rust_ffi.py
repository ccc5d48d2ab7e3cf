#!/usr/bin/env python
"""Generates bindings/folddisco_b200_sys.rs -- the Rust `extern "C"` block and #[repr(C)] structs of
include/folddisco_b200.h and include/folddisco_b200_host.h -- FROM THE HEADERS, so that the reference-side binding
shown in INTEGRATION.md cannot drift from the C ABI (round 1's hand-written block had lost two fields).

    python tools/gen_rust_ffi.py            # rewrite bindings/folddisco_b200_sys.rs
    python tools/gen_rust_ffi.py --check    # exit 1 if the committed file differs (tests/test_abi.py runs this)

The headers use a small C subset (fixed-width integers, float/double/char/int, pointers, fixed arrays, opaque
handles), which this script parses with regular expressions; anything it cannot map is an error, not a guess.
No Rust toolchain exists in the build image, so the output is checked structurally (every declared symbol and struct
appears once; sizes are asserted on the C side by tests/abi_smoke.c), not compiled.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = ["include/folddisco_b200.h", "include/folddisco_b200_host.h"]
OUT = os.path.join(ROOT, "bindings", "folddisco_b200_sys.rs")
SCALARS = {"uint64_t": "u64", "uint32_t": "u32", "uint16_t": "u16", "uint8_t": "u8", "int64_t": "i64", "int32_t": "i32",
           "int": "c_int", "float": "f32", "double": "f64", "char": "c_char", "void": "c_void", "size_t": "usize"}


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def rust_type(ctype, known):
    """C declarator type (without the name) -> Rust"""
    t = " ".join(ctype.replace("*", " * ").split())
    toks = t.split()
    base, i, const = None, 0, False
    while i < len(toks) and toks[i] != "*":
        if toks[i] == "const":
            const = True
        elif toks[i] in ("struct", "unsigned"):
            pass
        else:
            base = toks[i]
        i += 1
    if base in SCALARS:
        r = SCALARS[base]
    elif base in known:
        r = base
    else:
        raise SystemExit("gen_rust_ffi: unknown C type %r in %r" % (base, ctype))
    # pointers, left to right; a `const` after a star qualifies that pointer level
    while i < len(toks):
        assert toks[i] == "*", ctype
        r = ("*const " if const else "*mut ") + r
        const = False
        i += 1
        while i < len(toks) and toks[i] == "const":
            const = True
            i += 1
    return r


def split_decl(decl):
    """'const float *ca_xyz' -> ('const float *', 'ca_xyz', None); 'float U[9]' -> ('float', 'U', 9)"""
    decl = decl.strip()
    m = re.match(r"^(.*?)(\w+)\s*(?:\[(\d+)\])?$", decl, flags=re.S)
    if not m:
        raise SystemExit("gen_rust_ffi: cannot parse %r" % decl)
    return m.group(1).strip(), m.group(2), int(m.group(3)) if m.group(3) else None


def parse(text):
    text = strip_comments(text)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    text = text.replace('extern "C" {', " ").replace("}\n#endif", " ")
    opaque = re.findall(r"typedef\s+struct\s+(\w+)\s+\1\s*;", text)
    structs = []
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for stmt in m.group(1).split(";"):
            stmt = " ".join(stmt.split())
            if not stmt:
                continue
            # 'const float *n_xyz, *ca_xyz' style lists share the base type
            first, *rest = [p.strip() for p in stmt.split(",")]
            ctype, name, arr = split_decl(first)
            fields.append((ctype, name, arr))
            base = ctype.replace("*", "").strip()
            for r in rest:
                stars = r.count("*")
                c2, n2, a2 = split_decl(r.replace("*", " "))
                fields.append((base + " " + "*" * stars, n2, a2))
        structs.append((m.group(2), fields))
    body = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", text, flags=re.S)
    body = re.sub(r"typedef\s+struct\s+\w+\s+\w+\s*;", " ", body)
    aliases = re.findall(r"typedef\s+(fdh?_\w+)\s+(fdh?_\w+)\s*;", body)  # typedef fd_struct_row fdh_struct_row;
    body = re.sub(r"typedef\s+fdh?_\w+\s+fdh?_\w+\s*;", " ", body)
    funcs = []
    for m in re.finditer(r"([\w\s\*]+?)\b(fdh?_\w+)\s*\(([^()]*)\)\s*;", " ".join(body.split())):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        ps = []
        if params and params != "void":
            for k, p in enumerate(params.split(",")):
                ctype, pname, arr = split_decl(p)
                if not ctype:  # unnamed parameter
                    ctype, pname = p.strip(), "arg%d" % k
                ps.append((ctype, pname))
        funcs.append((ret, name, ps))
    return opaque, structs, funcs, aliases


def generate():
    opaque, structs, funcs, aliases = [], [], [], []
    for h in HEADERS:
        o, s, f, a = parse(open(os.path.join(ROOT, h)).read())
        opaque += o
        structs += s
        funcs += f
        aliases += a
    known = set(opaque) | {n for n, _ in structs} | {b for _, b in aliases}
    out = ["// GENERATED by tools/gen_rust_ffi.py from include/folddisco_b200.h and include/folddisco_b200_host.h.",
           "// Do not edit: re-run the script.  Analogue of lib/foldcomp/bindings.rs in the reference tree; see",
           "// INTEGRATION.md for the build.rs lines and the call sites this replaces.",
           "#![allow(non_camel_case_types, dead_code)]",
           "use std::os::raw::{c_char, c_int, c_void};", ""]
    for n in opaque:
        out += ["#[repr(C)] pub struct %s { _private: [u8; 0] }" % n]
    out.append("")
    for name, fields in structs:
        out.append("#[repr(C)]")
        out.append("#[derive(Clone, Copy)]")
        out.append("pub struct %s {" % name)
        for ctype, fname, arr in fields:
            rt = rust_type(ctype, known)
            if arr:
                rt = "[%s; %d]" % (rt, arr)
            out.append("    pub %s: %s," % ("r#type" if fname == "type" else fname, rt))
        out.append("}")
        out.append("")
    for a, b in aliases:
        out.append("pub type %s = %s;" % (b, a))
    if aliases:
        out.append("")
    consts = []
    for h in HEADERS:
        for m in re.finditer(r"#define\s+(FD_\w+)\s+\(?(-?\d+)\)?\s", open(os.path.join(ROOT, h)).read()):
            consts.append("pub const %s: c_int = %s;" % (m.group(1), m.group(2)))
    out += consts + [""]
    out.append('#[link(name = "folddisco_b200")]')
    out.append('extern "C" {')
    seen = set()
    for ret, name, ps in funcs:
        if name in seen:
            raise SystemExit("gen_rust_ffi: %s declared twice" % name)
        seen.add(name)
        args = ", ".join("%s: %s" % (("r#%s" % p) if p in ("type", "in", "ref", "move") else p, rust_type(c, known))
                         for c, p in ps)
        rr = "" if ret == "void" else " -> %s" % rust_type(ret, known)
        out.append("    pub fn %s(%s)%s;" % (name, args, rr))
    out.append("}")
    return "\n".join(out) + "\n", seen, {n for n, _ in structs}


def main():
    text, funcs, structs = generate()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            sys.stderr.write("bindings/folddisco_b200_sys.rs is stale: run python tools/gen_rust_ffi.py\n")
            sys.exit(1)
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(text)
    print("%s: %d functions, %d structs" % (os.path.relpath(OUT, ROOT), len(funcs), len(structs)))


if __name__ == "__main__":
    main()
