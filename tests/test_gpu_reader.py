"""Row f-4: the device catalogue reader (through the C ABI) against np.loadtxt -- the function the
reference's scripts call (/root/reference/tests/correlations.py:29-31) -- bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reader():
    from jax_powspec_b200 import reader
    return reader


def _write_catalogue(path, n, seed, trailing_newline=True, crlf=False, ncols=4):
    rng = np.random.default_rng(seed)
    v = rng.uniform(-20.0, 2520.0, (n, ncols))
    fmts = ["%.6f", "%.10e", "%.18e", "%.3f", "%g", "%.9g"]
    lines = []
    for i in range(n):
        f = fmts[i % len(fmts)]
        sep = (" ", "  ", "\t", " \t")[i % 4]
        line = sep.join(f % x for x in v[i])
        if i % 9 == 0:
            line = " " + line + "  "
        if i % 17 == 0:
            line += " # note"
        lines.append(line)
        if i % 101 == 0:
            lines.append("# comment only")
        if i % 203 == 0:
            lines.append("")
    eol = "\r\n" if crlf else "\n"
    text = eol.join(lines) + (eol if trailing_newline else "")
    with open(path, "w", newline="") as f:
        f.write(text)


def _same_bits(got, want):
    got = got.cpu().numpy()
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("trailing_newline,crlf", [(True, False), (False, False), (True, True)])
def test_matches_loadtxt(reader, tmp_path, trailing_newline, crlf):
    path = tmp_path / "cat.dat"
    _write_catalogue(path, 20000, 1, trailing_newline, crlf)
    want = np.loadtxt(path, usecols=(0, 1, 2), dtype=np.float32)
    got, info = reader.read_catalog_text(str(path), usecols=(0, 1, 2), return_info=True)
    _same_bits(got, want)
    assert info["n_rows"] == 20000 and info["n_host_rows"] == 0
    # column subset in another order, and the painters' strided column views
    _same_bits(reader.read_catalog_text(str(path), usecols=(3, 1)), np.loadtxt(path, usecols=(3, 1), dtype=np.float32))
    assert got[:, 1].stride(0) == 3


def test_box_mask_and_skiprows(reader, tmp_path):
    """particles[((particles < box) & (particles > 0)).all(axis=1)] of tests/correlations.py:30."""
    path = tmp_path / "cat.dat"
    _write_catalogue(path, 30000, 2)
    box = 2500.0
    p = np.loadtxt(path, usecols=(0, 1, 2), dtype=np.float32)
    want = p[((p < box) & (p > 0)).all(axis=1)]
    assert 0 < len(want) < len(p)
    _same_bits(reader.read_catalog_text(str(path), usecols=(0, 1, 2), box_size=box), want)
    want = np.loadtxt(path, usecols=(0, 1, 2), dtype=np.float32, skiprows=5)
    _same_bits(reader.read_catalog_text(str(path), usecols=(0, 1, 2), skiprows=5), want)


def test_host_deferred_fields_and_errors(reader, tmp_path):
    path = tmp_path / "odd.dat"
    rows = ["1.5 2.5 3.5", "nan 1 2", "1.000000000000000000000001 2 3", "4 inf 5", "1e-40 2 3", "7 8 9",
            "12345678901234567891 1 1"]
    path.write_text("\n".join(rows) + "\n")
    want = np.loadtxt(path, usecols=(0, 1, 2), dtype=np.float32)
    got, info = reader.read_catalog_text(str(path), return_info=True)
    assert info["n_host_rows"] == 5
    g = got.cpu().numpy()
    assert np.array_equal(np.isnan(g), np.isnan(want))
    assert np.array_equal(g[~np.isnan(g)].view(np.uint32), want[~np.isnan(want)].view(np.uint32))
    # with the box mask the deferred rows are masked on the host with the same rule
    m = ((want < 100) & (want > 0)).all(axis=1)
    _same_bits(reader.read_catalog_text(str(path), box_size=100.0), want[m])
    bad = tmp_path / "bad.dat"
    bad.write_text("1 2 3\n4 five 6\n7 8 9\n1 2\n")
    with pytest.raises(ValueError, match="line 2"):
        reader.read_catalog_text(str(bad))
    short = tmp_path / "short.dat"
    short.write_text("1 2 3\n4 5\n")
    with pytest.raises(ValueError):
        reader.read_catalog_text(str(short))
    empty = tmp_path / "empty.dat"
    empty.write_text("")
    assert reader.read_catalog_text(str(empty)).shape == (0, 3)
    only_comments = tmp_path / "c.dat"
    only_comments.write_text("# a\n\n# b\n")
    assert reader.read_catalog_text(str(only_comments)).shape == (0, 3)


def test_large_file_feeds_the_painter(reader, tmp_path):
    """5e5 rows (chunk boundaries inside numbers, > 1 scan block) and straight into cic_mas_vec."""
    import jax_powspec_b200 as jps
    rng = np.random.default_rng(5)
    n = 500_000
    p = rng.uniform(0, 1000, (n, 3))
    path = tmp_path / "big.dat"
    np.savetxt(path, p, fmt="%.8f")
    want = np.loadtxt(path, dtype=np.float32)
    got = reader.read_catalog_text(str(path), box_size=1000.0)
    keep = ((want < 1000.0) & (want > 0)).all(axis=1)
    _same_bits(got, want[keep])
    nm = 64
    mesh = jps.cic_mas_vec(torch.zeros((nm, nm, nm), device="cuda"), got[:, 0], got[:, 1], got[:, 2], None, len(got),
                           0.0, 0.0, 0.0, 1000.0, nm, True)
    assert abs(float(mesh.sum()) - len(got)) < 1e-3 * len(got)
