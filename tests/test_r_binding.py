"""The R binding as files (r/): the `.Call` shim r/src/mcstate_b200_r.c and the R6
`dust_generator` r/R/generator.R -- what binds libmcstate_b200.so into mcstate
(R/particle_filter.R:590-593 is_dust_generator; R/particle_filter_state.R:237-263 the calls;
SURVEY.md 8(b) the method table).  R is not in the image, so the files cannot be loaded; what
CAN be checked is checked here:

  * the C file type-checks (gcc -fsyntax-only) against the real include/mcstate_b200.h and
    stub headers of R's C API (r/stub/): every mcb_* call has the right arguments;
  * every entry of its R_CallMethodDef table names a defined mcbr_* function of that arity;
  * every mcb_* function it calls is declared in the header and exported by the library;
  * every .Call("mcbr_*", ...) in generator.R names a registered routine with the right
    number of arguments;
  * the R6 class has every method of SURVEY.md 8(b)'s table and the generator attribute.
"""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "r", "src", "mcstate_b200_r.c")
GEN = os.path.join(ROOT, "r", "R", "generator.R")
HEADER = os.path.join(ROOT, "include", "mcstate_b200.h")


def _strip_c_comments(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def _split_args(s):
    """Top-level comma split of an argument list (parentheses, brackets and strings nest)."""
    out, depth, cur, quote = [], 0, "", None
    for ch in s:
        if quote:
            cur += ch
            if ch == quote:
                quote = None
            continue
        if ch in "\"'":
            quote = ch
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _call_args(text, start):
    """The argument string of the call whose '(' is at text[start]."""
    depth, i, quote = 0, start, None
    while True:
        ch = text[i]
        if quote:
            if ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return text[start + 1:i]
        i += 1


def _table():
    src = _strip_c_comments(open(SHIM).read())
    body = src[src.index("R_CallMethodDef calls[]"):]
    return {m.group(1): int(m.group(3)) for m in
            re.finditer(r'\{"(mcbr_\w+)",\s*\(DL_FUNC\)&(\w+),\s*(\d+)\}', body)
            if m.group(1) == m.group(2)}


def test_shim_type_checks_against_the_header():
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                        "-Wno-cast-function-type", "-I", os.path.join(ROOT, "r", "stub"),
                        "-I", os.path.join(ROOT, "include"), SHIM], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_registration_table_matches_the_definitions():
    src = _strip_c_comments(open(SHIM).read())
    table = _table()
    assert len(table) >= 25
    defined = {}
    for m in re.finditer(r"^SEXP (mcbr_\w+)\(", src, flags=re.M):
        args = _call_args(src, m.end() - 1).strip()
        defined[m.group(1)] = 0 if args == "void" else len(_split_args(args))
        if args != "void":
            assert all(a.startswith("SEXP ") for a in _split_args(args)), m.group(1)
    assert defined == table          # every routine registered, every arity right
    assert "R_init_mcstateb200" in src and "R_registerRoutines" in src


def test_every_engine_call_is_declared_and_exported():
    from mcstate_b200 import _lib

    src = _strip_c_comments(open(SHIM).read())
    used = set(re.findall(r"\b(mcb_[a-z_0-9]+)\(", src))
    header = _strip_c_comments(open(HEADER).read())
    declared = set(re.findall(r"\b(mcb_[a-z_0-9]+)\(", header))
    assert used and used <= declared, used - declared
    lib = C.CDLL(_lib.LIB_PATH)
    for name in used:
        assert hasattr(lib, name), name
    # the methods of SURVEY.md 8(b)'s table all reach the engine
    for name in ("mcb_alloc", "mcb_free", "mcb_run", "mcb_filter", "mcb_set_data", "mcb_update_state",
                 "mcb_state", "mcb_reorder", "mcb_set_index", "mcb_rng_state", "mcb_set_rng_state",
                 "mcb_time", "mcb_simulate", "mcb_compare_data", "mcb_resample",
                 "mcb_rng_distributed_state", "mcb_last_error"):
        assert name in used, name


def test_generator_calls_registered_routines_with_the_right_arity():
    table = _table()
    text = open(GEN).read()
    seen = set()
    for m in re.finditer(r'\.Call\(', text):
        args = _split_args(_call_args(text, m.end() - 1))
        name = args[0].strip('"')
        assert args[-1].replace(" ", "") == 'PACKAGE="mcstateb200"', args
        assert name in table, name
        assert len(args) - 2 == table[name], (name, len(args) - 2, table[name])
        seen.add(name)
    assert seen == set(table)        # nothing registered that the R side never calls


def test_generator_has_the_interface_mcstate_drives():
    text = open(GEN).read()
    assert 'attr(gen, "name") <- "dust_generator"' in text      # R/particle_filter.R:590-593
    methods = set(re.findall(r"^\s{6}(\w+) = function\(", text, flags=re.M))
    want = {"initialize", "run", "filter", "set_data", "update_state", "state", "reorder",
            "set_index", "rng_state", "set_rng_state", "info", "pars", "shape", "n_state", "n_pars",
            "time", "n_threads", "set_n_threads", "simulate", "uses_gpu", "compare_data", "resample",
            "has_compare", "has_gpu_support", "has_openmp", "time_type", "gpu_info", "name",
            "n_particles"}
    assert want <= methods, want - methods
    # $new's arguments as R/particle_filter_state.R:237-241 passes them
    init = re.search(r"initialize = function\(([^)]*)\)", text).group(1)
    for arg in ("pars", "time", "n_particles", "n_threads", "seed", "pars_multi", "deterministic",
                "gpu_config"):
        assert re.search(r"\b%s\b" % arg, init), arg
    # no elided bodies
    assert "..." not in re.sub(r"#.*", "", text)
    # the package files a maintainer needs
    for f in ("DESCRIPTION", "NAMESPACE", os.path.join("src", "Makevars")):
        assert os.path.exists(os.path.join(ROOT, "r", f)), f
