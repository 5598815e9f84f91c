"""Row f-4 (catalogue text reader): the decimal -> float32 converter the device runs
(jax_powspec_b200/csrc/textparse.cuh), compiled for the host and fuzzed against Python's float()
(= strtod, what np.loadtxt uses: /root/reference/tests/correlations.py:29) -- bit-exact or
explicitly deferred, never approximate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "helpers", "textparse_host.cpp")
OUT = os.path.join(HERE, "helpers", "_build", "libtextparse_host.so")
OK, SLOW, BAD, MISSING = 0, 1, 2, 3


@pytest.fixture(scope="module")
def host():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(HERE, "..", "jax_powspec_b200", "csrc", "textparse.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    lib = C.CDLL(OUT)
    lib.jps_host_parse_fields.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.jps_host_parse_fields.restype = None
    lib.jps_host_parse_line.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int,
                                        C.POINTER(C.c_float), C.POINTER(C.c_int)]
    return lib


def parse_fields(lib, toks):
    enc = [t.encode() for t in toks]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(e) for e in enc])
    buf = b"".join(enc)
    out = np.zeros(len(enc), dtype=np.float32)
    st = np.zeros(len(enc), dtype=np.int32)
    lib.jps_host_parse_fields(buf, off.ctypes.data, len(enc), out.ctypes.data, st.ctypes.data)
    return out, st


def expected(toks):
    return np.array([float(t) for t in toks], dtype=np.float64).astype(np.float32)


def check_exact(lib, toks, allow_slow=False):
    out, st = parse_fields(lib, toks)
    want = expected(toks)
    if not allow_slow:
        bad = np.nonzero(st != OK)[0]
        assert bad.size == 0, (toks[bad[0]], st[bad[0]])
    ok = st == OK
    assert set(np.unique(st)) <= {OK, SLOW}
    a, b = out[ok].view(np.uint32), want[ok].view(np.uint32)
    diff = np.nonzero(a != b)[0]
    assert diff.size == 0, (np.array(toks)[ok][diff[0]], out[ok][diff[0]], want[ok][diff[0]])
    return st


@pytest.mark.parametrize("fmt", ["%.6f", "%.3f", "%.8e", "%.18e", "%.17g", "%g", "%.9g", "%.15e", "%.1f", "%d"])
def test_formats_bit_exact(host, fmt):
    rng = np.random.default_rng(hash(fmt) % 2**32)
    vals = np.concatenate([rng.uniform(0, 2500, 40000), rng.uniform(-1000, 1000, 20000),
                           10.0 ** rng.uniform(-8, 12, 20000) * rng.choice([-1, 1], 20000),
                           rng.integers(0, 10**9, 5000).astype(np.float64)])
    if fmt == "%d":
        toks = [fmt % int(v) for v in vals]
    else:
        toks = [fmt % v for v in vals]
    check_exact(host, toks)


def test_float32_ties_and_halfway_doubles(host):
    """Decimal expansions of exact float32 midpoints and of neighbours of double midpoints: the
    double step must round correctly or the float32 result moves by one ulp."""
    rng = np.random.default_rng(7)
    f = rng.uniform(1.0, 4096.0, 20000).astype(np.float32)
    nxt = np.nextafter(f, np.float32(np.inf))
    mid = (f.astype(np.float64) + nxt.astype(np.float64)) / 2          # exact in float64
    toks = []
    for m in mid:
        s = "%.40f" % m                                                # exact decimal expansion of the midpoint
        s = s.rstrip("0")
        toks.append(s)
    st = check_exact(host, toks, allow_slow=True)
    assert (st == OK).mean() > 0.9                                     # <= 19 significant digits most of the time
    # one unit in the last kept place above / below the midpoint
    up = [t[:-1] + str(int(t[-1]) + 1) if t[-1] != "9" else t + "1" for t in toks]
    check_exact(host, up, allow_slow=True)


def test_nineteen_digits_and_exponent_envelope(host):
    rng = np.random.default_rng(11)
    toks = []
    for _ in range(30000):
        nd = int(rng.integers(1, 20))
        digits = str(int(rng.integers(1, 10))) + "".join(str(int(d)) for d in rng.integers(0, 10, nd - 1))
        e = int(rng.integers(-27, 28)) - (nd - 1)
        toks.append(f"{digits[0]}.{digits[1:]}e{e + nd - 1:+d}" if nd > 1 else f"{digits}e{e:+d}")
    check_exact(host, toks, allow_slow=True)
    out, st = parse_fields(host, toks)
    assert (st == OK).mean() > 0.4
    # plain integers up to 2^64, leading zeros, bare dots, signs
    special = ["0", "-0", "+0.0", "000123.4500", ".5", "5.", "-.25", "+1e3", "1E3", "1e+03", "1e-03",
               "18446744073709551615", "9007199254740993", "9007199254740992", "4503599627370497.5",
               "0.000000000000000000000000001", "1e27", "1e28", "123456789012345678900000", "1e30", "0e999"]
    st = check_exact(host, special, allow_slow=True)
    assert st[special.index("1e28")] == OK and st[special.index("1e30")] == OK    # normalised into the envelope


def test_deferred_and_rejected_fields(host):
    out, st = parse_fields(host, ["nan", "inf", "-inf", "NaN", "Infinity", "1e400", "1e-400",
                                  "1.00000000000000000001", "12345678901234567891"])
    assert (st == SLOW).all()
    out, st = parse_fields(host, ["abc", "1.2.3", "1e", "1e+", "--1", "1,5", "0x10", "1_000", "+", "-", ".", "e5", "1.5f"])
    assert (st == BAD).all(), st
    # digits beyond 19 that are all zero are exact
    out, st = parse_fields(host, ["1.0000000000000000000000", "250000000000000000000000"])
    assert (st == OK).all() and out[0] == 1.0 and out[1] == np.float32(2.5e23)


def test_lines_match_loadtxt(host, tmp_path):
    rng = np.random.default_rng(3)
    rows = []
    for i in range(1500):
        v = rng.uniform(-10, 2510, 5)
        sep = rng.choice([" ", "  ", "\t", " \t "])
        line = sep.join(("%.6f" if i % 3 else "%.10e") % x for x in v)
        if i % 7 == 0:
            line = "  " + line + "   "
        if i % 11 == 0:
            line += "  # trailing comment"
        if i % 13 == 0:
            line += "\r"
        rows.append(line)
        if i % 50 == 0:
            rows.append("# a comment line")
        if i % 70 == 0:
            rows.append("")
            rows.append("   \t ")
    path = tmp_path / "cat.txt"
    path.write_text("\n".join(rows) + "\n")
    want = np.loadtxt(path, usecols=(0, 1, 2), dtype=np.float32)
    want41 = np.loadtxt(path, usecols=(4, 1), dtype=np.float32)
    got, got41 = [], []
    for cols, dst in (((0, 1, 2), got), ((4, 1), got41)):
        cc = (C.c_int * len(cols))(*cols)
        for line in rows:
            b = line.encode()
            vals = (C.c_float * 8)()
            status = C.c_int(0)
            if host.jps_host_parse_line(b, len(b), ord("#"), cc, len(cols), vals, C.byref(status)):
                assert status.value == OK, line
                dst.append([vals[k] for k in range(len(cols))])
    assert np.array_equal(np.array(got, dtype=np.float32).view(np.uint32), want.view(np.uint32))
    assert np.array_equal(np.array(got41, dtype=np.float32).view(np.uint32), want41.view(np.uint32))
    # a short row is an error, as in np.loadtxt
    b = b"1.0 2.0"
    cc = (C.c_int * 3)(0, 1, 2)
    vals = (C.c_float * 8)()
    status = C.c_int(0)
    assert host.jps_host_parse_line(b, len(b), ord("#"), cc, 3, vals, C.byref(status)) == 1
    assert status.value == MISSING
