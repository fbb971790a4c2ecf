"""CPU: the Lua-free parameter-file reader (pmc_rc.cpp, readConf API of pmctools/include/readConf.h:44-89) and the pieces of
the pmc_rc layer that need no GPU; plus the reference's own vanilla_pmc binary (when it was built here) failing loudly without
a device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pmclib_b200.lib import LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp = C.c_void_p


@pytest.fixture(scope="module")
def L():
    lib = C.CDLL(LIB_PATH)
    lib.initError.restype = vp
    lib.rc_open.restype = vp; lib.rc_open.argtypes = [C.c_char_p, vp]
    lib.rc_alias.restype = vp; lib.rc_alias.argtypes = [vp, C.c_char_p, vp]
    lib.rc_close.argtypes = [vp]
    lib.rc_do_args.argtypes = [vp, C.c_int, vp, C.c_char, C.c_char_p, vp]
    lib.rc_has_key.argtypes = [vp, C.c_char_p, vp]
    lib.rc_is_array.argtypes = [vp, C.c_char_p, vp]
    lib.rc_get_real.restype = C.c_double; lib.rc_get_real.argtypes = [vp, C.c_char_p, vp]
    lib.rc_get_integer.restype = C.c_long; lib.rc_get_integer.argtypes = [vp, C.c_char_p, vp]
    lib.rc_get_string.restype = C.c_char_p; lib.rc_get_string.argtypes = [vp, C.c_char_p, vp]
    lib.rc_safeget_integer.restype = C.c_long; lib.rc_safeget_integer.argtypes = [vp, C.c_char_p, C.c_long, vp]
    lib.rc_array_size.restype = C.c_size_t; lib.rc_array_size.argtypes = [vp, C.c_char_p, vp]
    lib.rc_get_real_array.restype = C.c_size_t; lib.rc_get_real_array.argtypes = [vp, C.c_char_p, vp, vp]
    lib.rc_get_integer_array.restype = C.c_size_t; lib.rc_get_integer_array.argtypes = [vp, C.c_char_p, vp, vp]
    lib.rc_get_string_array.restype = C.c_size_t; lib.rc_get_string_array.argtypes = [vp, C.c_char_p, vp, vp]
    lib.rc_get_real_assquarematrix.restype = C.c_size_t; lib.rc_get_real_assquarematrix.argtypes = [vp, C.c_char_p, vp, C.c_int, vp]
    lib.get_itername.restype = C.c_char_p; lib.get_itername.argtypes = [vp, C.c_char_p, C.c_int, vp]
    lib.getErrorValue.argtypes = [vp]; lib._isError.argtypes = [vp]; lib.purgeError.argtypes = [vp]
    lib.parabox_from_rc.restype = vp; lib.parabox_from_rc.argtypes = [vp, C.c_char_p, vp]
    lib.rc_gsl_rng.restype = vp; lib.rc_gsl_rng.argtypes = [vp, C.c_char_p, vp]
    lib.init_distribution_from_rc.restype = vp; lib.init_distribution_from_rc.argtypes = [vp, C.c_char_p, vp]
    return lib


def _err(L):
    e = vp(L.initError())
    return e, C.byref(e)


def _reals(L, rc, key, perr):
    p = C.POINTER(C.c_double)()
    n = L.rc_get_real_array(rc, key, C.byref(p), perr)
    return [p[i] for i in range(n)]


def test_parameter_file_subset(L, tmp_path):
    f = tmp_path / "a.par"
    f.write_text('''
-- a comment
a = {}
a.x = 3          a.y = -.5e1 ; a.s = "str" .. 'ing'
a.list = {1, 2, 3.5,}
a.nested = { deep = { v = 2^3 + 1 }, names = {"u", "v"}, [3] = 7 }
--[[ block
comment ]]
b = a              -- reference, not copy
b.z = (1 + 2) * 4 / 3
m = { {1, 2}, {3, 4} }
flag = true
local q = 5
fmt = "out_%d.dat"
files = { [0] = "zero", [1] = "one" }
''')
    e, pe = _err(L)
    rc = L.rc_open(str(f).encode(), pe)
    assert rc and not L._isError(e)
    assert L.rc_get_integer(rc, b"a.x", pe) == 3 and L.rc_get_real(rc, b"a.y", pe) == -5.0
    assert L.rc_get_string(rc, b"a.s", pe) == b"string"
    assert _reals(L, rc, b"a.list", pe) == [1.0, 2.0, 3.5] and L.rc_array_size(rc, b"a.list", pe) == 3
    assert L.rc_get_real(rc, b"a.nested.deep.v", pe) == 9.0 and L.rc_get_integer(rc, b"a.nested[3]", pe) == 7
    assert L.rc_get_real(rc, b"a.z", pe) == 4.0          # b = a shares the table, as in Lua
    assert L.rc_get_real(rc, b"m[2][1]", pe) == 3.0 and L.rc_get_integer(rc, b"q", pe) == 5
    assert L.rc_has_key(rc, b"a.nope", pe) == 0 and L.rc_has_key(rc, b"a.nested.deep", pe) == 1
    assert L.rc_is_array(rc, b"a.list", pe) == 1 and L.rc_is_array(rc, b"a.x", pe) == 0
    assert L.rc_safeget_integer(rc, b"a.nope", 42, pe) == 42
    sp = C.POINTER(C.c_char_p)()
    assert L.rc_get_string_array(rc, b"a.nested.names", C.byref(sp), pe) == 2 and sp[0] == b"u" and sp[1] == b"v"
    # aliases: keys relative to a prefix, nested aliases, index after a prefix
    al = vp(L.rc_alias(rc, b"a.nested", pe))
    assert L.rc_get_real(al, b"deep.v", pe) == 9.0
    al2 = vp(L.rc_alias(al, b"deep", pe))
    assert L.rc_get_real(al2, b"v", pe) == 9.0
    L.rc_close(C.byref(al2)); L.rc_close(C.byref(al))
    # square matrices: scalar, diagonal, flat, rows (readConf.c:697-760)
    mp = C.POINTER(C.c_double)()
    assert L.rc_get_real_assquarematrix(rc, b"a.x", C.byref(mp), 2, pe) == 4 and [mp[i] for i in range(4)] == [3, 0, 0, 3]
    assert L.rc_get_real_assquarematrix(rc, b"m", C.byref(mp), 2, pe) == 4 and [mp[i] for i in range(4)] == [1, 2, 3, 4]
    assert L.rc_get_real_assquarematrix(rc, b"a.nested.names", C.byref(mp), 2, pe) == 0 and L._isError(e)
    L.purgeError(pe)
    # get_itername: printf pattern or array indexed by the iteration (pmc_rc.c:770-794)
    assert L.get_itername(rc, b"fmt", 3, pe) == b"out_3.dat" and L.get_itername(rc, b"files", 1, pe) == b"one"
    # missing key and wrong type raise rc_bad_key / rc_bad_type
    L.rc_get_real(rc, b"a.nope", pe)
    assert L.getErrorValue(e) == -205
    L.purgeError(pe)
    L.rc_get_real(rc, b"a.s", pe)
    assert L.getErrorValue(e) == -204
    L.purgeError(pe)
    # -e overrides (readConf.c:982-1060)
    argv = (C.c_char_p * 4)(b"prog", b"-e", b"a.x = 11  extra = {1,2}", str(f).encode())
    L.rc_do_args(rc, 4, argv, b"e", b"extra", pe)
    assert L.rc_get_integer(rc, b"a.x", pe) == 11 and L.rc_array_size(rc, b"extra", pe) == 2
    L.rc_close(C.byref(vp(rc)))


def test_syntax_errors_are_reported(L, tmp_path):
    e, pe = _err(L)
    for text in ("a = = 3", "a = {1, 2", "f(3)", "a = b.c.d", 'a = "open'):
        f = tmp_path / "bad.par"
        f.write_text(text)
        assert not L.rc_open(str(f).encode(), pe)
        assert L.getErrorValue(e) == -202
        L.purgeError(pe)


def test_host_side_objects_from_a_parameter_file(L, tmp_path):
    """parabox_from_rc, rc_gsl_rng (MT19937 with GSL's seeding: the first outputs of gsl-1.14's mt19937 for seed 4357 and
    for seed 1), init_distribution_from_rc with conditional parameters."""
    f = tmp_path / "b.par"
    f.write_text('box = { min = {-1, -2, -3}, max = {1, 2, 3} }\nr1 = { seed = 1 }\nr0 = {}\n'
                 't = { name = "bentGauss", ndim = 4, bent = 0.1, sig1_square = 10, conditional = { id = {2, 3}, value = {1, 2} } }\n')
    e, pe = _err(L)
    rc = L.rc_open(str(f).encode(), pe)

    class Pb(C.Structure):
        _fields_ = [("ndim", C.c_int), ("nfaces", C.c_int), ("faces", C.POINTER(C.c_double))]
    pb = C.cast(L.parabox_from_rc(rc, b"box", pe), C.POINTER(Pb)).contents
    assert pb.ndim == 3 and pb.nfaces == 6
    faces = np.array([pb.faces[i] for i in range(6 * 4)]).reshape(6, 4)
    assert np.array_equal(faces[0], [-1, 0, 0, 1]) and np.array_equal(faces[5], [0, 0, 1, 3])

    class RngType(C.Structure):
        _fields_ = [("name", C.c_char_p), ("max", C.c_ulong), ("min", C.c_ulong), ("size", C.c_size_t), ("set", vp),
                    ("get", C.CFUNCTYPE(C.c_ulong, vp)), ("get_double", C.CFUNCTYPE(C.c_double, vp))]

    class Rng(C.Structure):
        _fields_ = [("type", C.POINTER(RngType)), ("state", vp)]
    r0 = C.cast(L.rc_gsl_rng(rc, b"r0", pe), C.POINTER(Rng)).contents
    assert r0.type.contents.name == b"mt19937"
    assert [r0.type.contents.get(r0.state) for _ in range(3)] == [4293858116, 699692587, 1213834231]  # gsl_rng_default_seed 0 -> 4357
    r1 = C.cast(L.rc_gsl_rng(rc, b"r1", pe), C.POINTER(Rng)).contents
    assert [r1.type.contents.get(r1.state) for _ in range(3)] == [1791095845, 4282876139, 3093770124]  # MT19937, seed 1

    class Dist(C.Structure):
        _fields_ = [("ndim", C.c_int), ("n_ded", C.c_int), ("ndef", C.c_int), ("pars", C.POINTER(C.c_double)), ("def_", C.POINTER(C.c_int))]
    t = C.cast(L.init_distribution_from_rc(rc, b"t", pe), C.POINTER(Dist)).contents
    assert not L._isError(e)
    assert t.ndim == 2 and t.ndef == 2 and [t.def_[i] for i in range(4)] == [0, 0, 1, 1] and [t.pars[i] for i in range(4)][2:] == [1.0, 2.0]


def test_reference_vanilla_pmc_fails_loudly_without_a_device(tmp_path):
    exe = os.path.join(ROOT, "oracle", "_ref", "vanilla_pmc")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/vanilla_pmc not built")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_gpu_rc.py runs the driver")
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "c", "test_pmc.par")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "-6024" in r.stderr and "no CPU fallback" in r.stderr
