#!/usr/bin/env python
"""Generates fortran/swcuda_c_binding.f90 -- the iso_c_binding interface module of libswcuda.so -- from
include/swcuda.h, so that the Fortran side declares EVERY entry point of the C ABI with the header's own
argument lists (tests/test_shim.py regenerates it and compares).

    python fortran/gen_bindings.py            # rewrite fortran/swcuda_c_binding.f90
    python fortran/gen_bindings.py --check    # exit 1 if the committed file is stale

Type map: int/long/double/unsigned long long by value; `const T *` / `T *` data pointers (host OR device
arrays, blobs) as type(c_ptr), value -- the caller passes c_loc(array) or c_devloc(array); pointers to the
ABI's structs by reference; the scalars a function returns through a pointer as intent(out) arguments."""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "swcuda.h")
OUT = os.path.join(HERE, "swcuda_c_binding.f90")

STRUCTS = {"swcu_dims", "swcu_params", "swh_basin"}
# scalars returned through a pointer: (function, argument) -> Fortran declaration
OUT_SCALARS = {
    ("swcu_synchronize", "bad_cells"): "integer(c_long), intent(out)",
    ("swcu_timer_stop", "elapsed_ms"): "real(c_float), intent(out)",
    ("swcu_march_band_rows", "first"): "integer(c_int), intent(out)",
    ("swcu_march_band_rows", "last"): "integer(c_int), intent(out)",
    ("swcu_halo_plan", "send_row"): "integer(c_int), intent(out)",
    ("swcu_halo_plan", "recv_row"): "integer(c_int), intent(out)",
    ("swh_uniform_split", "start"): "integer(c_int), intent(out)",
    ("swh_uniform_split", "size"): "integer(c_int), intent(out)",
    ("swh_hilbert_d2xy", "x"): "integer(c_int), intent(out)",
    ("swh_hilbert_d2xy", "y"): "integer(c_int), intent(out)",
    ("swcu_selftest_mdiv", "mismatches"): "integer(c_long), intent(out)",
    ("swcu_profile_steps", "prep_ms"): "real(c_float), intent(out)",
    ("swcu_profile_steps", "prep_launches"): "integer(c_long), intent(out)",
    ("swcu_profile_steps", "update_ms"): "real(c_float), intent(out)",
    ("swcu_profile_steps", "update_launches"): "integer(c_long), intent(out)",
}
SCALARS = {"int": "integer(c_int)", "long": "integer(c_long)", "double": "real(c_double)", "float": "real(c_float)",
           "unsigned long long": "integer(c_long_long)"}


def prototypes(text):
    """[(return type, name, [(type, name), ...])] of every function declared in the header."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    text = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", " ", text, flags=re.S)
    text = re.sub(r"enum\s+\w+\s*\{.*?\}\s*;", " ", text, flags=re.S)
    text = re.sub(r"typedef\s+struct\s+\w+\s+\w+\s*;", " ", text)
    text = text.replace('extern "C" {', " ")
    out = []
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(sw[ch]u?_\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        ret = " ".join(m.group(1).split())
        args = []
        body = " ".join(m.group(3).split())
        if body and body != "void":
            for a in body.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)(\[\d*\])?$", a)
                t, n, arr = mm.group(1).strip(), mm.group(2), mm.group(3)
                if arr:
                    t += " *"
                args.append((" ".join(t.split()), n))
        out.append((ret, m.group(2), args))
    return out


def fortran_arg(fn, ctype, name):
    base = ctype.replace("const", "").replace("*", "").strip()
    stars = ctype.count("*")
    if (fn, name) in OUT_SCALARS:
        return OUT_SCALARS[(fn, name)], set()
    if stars == 0:
        return SCALARS[base] + ", value", set()
    if base in STRUCTS and stars == 1:
        return f"type({base}), intent(in)", {base}
    if base == "swcu_ctx" and stars == 2 and "const" not in ctype:
        return "type(c_ptr), intent(out)", set()            # swcu_create(swcu_ctx **out, ...)
    if base == "swcu_ctx" and stars == 2:
        return "type(c_ptr), intent(in), dimension(*)", set()   # swcu_ctx *const *ctxs
    if base == "char":
        return "character(kind=c_char), intent(in), dimension(*)", set()
    return "type(c_ptr), value", set()                      # data pointer, context handle, stream


def fortran_ret(ctype):
    if "*" in ctype:
        return "type(c_ptr)"
    return SCALARS[ctype.replace("const", "").strip()]


def constants(text):
    """#define'd integers and the three enums of the header as Fortran parameters."""
    lines = []
    for m in re.finditer(r"^#define\s+(SWCU_\w+)\s+(-?\d+)\b", text, flags=re.M):
        lines.append(f"    integer(c_int), parameter :: {m.group(1)} = {m.group(2)}")
    for m in re.finditer(r"enum\s+(\w+)\s*\{(.*?)\}\s*;", text, flags=re.S):
        body = re.sub(r"/\*.*?\*/", " ", m.group(2), flags=re.S)
        val = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, v = [x.strip() for x in item.split("=")]
                val = int(v)
            else:
                name, val = item, val + 1
            lines.append(f"    integer(c_int), parameter :: {name} = {val}")
    lines.append("    integer(c_int), parameter :: SWCU_NF4 = SWCU_F4_END - 100")
    return lines


def generate():
    text = open(HEADER).read()
    out = ["!-----------------------------------------------------------------------------------------------",
           "! swcuda_c_binding.f90 -- GENERATED by fortran/gen_bindings.py from include/swcuda.h; do not edit.",
           "! bind(C) interfaces of every entry point of libswcuda.so, the constants and the three structs.",
           "!-----------------------------------------------------------------------------------------------",
           "module swcuda_c_binding", "    use iso_c_binding", "    implicit none", ""]
    out += constants(text)
    out += ["",
            "    type, bind(C) :: swcu_dims", "        integer(c_int) :: nx_start, nx_end, ny_start, ny_end",
            "        integer(c_int) :: bnd_x1, bnd_x2, bnd_y1, bnd_y2", "    end type", "",
            "    type, bind(C) :: swcu_params", "        integer(c_int) :: full_free_surface, trans_terms, ksw_lat",
            "        real(c_double) :: time_smooth", "        integer(c_int) :: use_tracers", "        integer(c_int) :: mode",
            "    end type", "",
            "    type, bind(C) :: swh_basin", "        integer(c_int) :: nx, ny", "        real(c_double) :: dxst, dyst, rlon, rlat",
            "        integer(c_int) :: curve_grid", "        real(c_double) :: rotation_on_lon, rotation_on_lat", "    end type", "",
            "    interface"]
    for ret, name, args in prototypes(text):
        names = [a[1] for a in args]
        kinds, structs, decls = {"c_ptr"}, set(), []
        for t, n in args:
            d, st = fortran_arg(name, t, n)
            structs |= st
            decls.append(f"            {d} :: {n}")
            kinds |= set(re.findall(r"c_\w+", d))
        r = fortran_ret(ret)
        kinds |= set(re.findall(r"c_\w+", r))
        head = f"        function {name}({', '.join(names)}) bind(C, name=\"{name}\") result(rc_)"
        if len(head) > 120:       # continuation lines
            parts, cur = [], f"        function {name}("
            for i, n in enumerate(names):
                piece = n + (", " if i + 1 < len(names) else ")")
                if len(cur) + len(piece) > 110:
                    parts.append(cur + "&")
                    cur = "                " + piece
                else:
                    cur += piece
            parts.append(cur + " &")
            parts.append(f"                bind(C, name=\"{name}\") result(rc_)")
            head = "\n".join(parts)
        out.append(head)
        out.append("            import :: " + ", ".join(sorted(kinds) + sorted(structs)))
        out += decls
        out.append(f"            {r} :: rc_")
        out.append("        end function")
    out += ["    end interface", "end module swcuda_c_binding", ""]
    return "\n".join(out)


if __name__ == "__main__":
    new = generate()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == new else 1)
    open(OUT, "w").write(new)
    print(OUT)
