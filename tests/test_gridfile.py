"""V3 "OMGRID" grid files: this repository's reader/writer against the reference's own
GridForce::saveToFile / GridData::saveToFile / GridForce::loadFromFile (through oracle/_ref where the reference compiles;
against committed bytes everywhere), mirroring the reference's round-trip tests (python/tests/test_auto_grid.py:54-102,
allclose rtol 1e-10 — here bit-exact)."""
import os

import numpy as np
import pytest

import openmmgridforce_b200 as gf

COUNTS, SPACING, ORIGIN = (5, 6, 7), (0.1, 0.125, 0.15), (1.00175115, 0.5328844699999999, 0.8606374500000002)


def _values():
    return np.random.default_rng(3).normal(size=COUNTS) * 41.84


def test_round_trip_bit_exact(tmp_path):
    path = str(tmp_path / "g.grid")
    v = _values()
    gf.write_grid_file(path, COUNTS, SPACING, ORIGIN, v, grid_type=1, inv_power=4.0, inv_power_mode=2, with_trailer=True)
    h, w = gf.read_grid_file(path)
    assert h["counts"] == COUNTS and h["spacing"] == SPACING and h["origin"] == ORIGIN
    assert (h["grid_type"], h["inv_power"], h["inv_power_mode"], h["deriv_count"], h["data_offset"]) == (1, 4.0, 2, 0, 128)
    assert np.array_equal(v, w)
    assert os.path.getsize(path) == 128 + 8 * v.size + 4 + 24


def test_known_file_bytes(tmp_path):
    """Header layout pinned field by field (Appendix A of SURVEY.md; GridForce.cpp:723-787)."""
    path = str(tmp_path / "g.grid")
    gf.write_grid_file(path, (2, 3, 4), (0.5, 0.25, 0.125), (1.0, 2.0, 3.0), np.arange(24.0), grid_type=3)
    raw = open(path, "rb").read()
    assert raw[:8] == b"OMGRID\0\0"
    assert np.frombuffer(raw, dtype="<u4", count=2, offset=8).tolist() == [3, 128]
    assert np.frombuffer(raw, dtype="<i4", count=4, offset=16).tolist() == [2, 3, 4, 0]
    assert np.frombuffer(raw, dtype="<f8", count=3, offset=32).tolist() == [0.5, 0.25, 0.125]
    assert np.frombuffer(raw, dtype="<u8", count=1, offset=56).tolist() == [128]
    assert np.frombuffer(raw, dtype="<f8", count=3, offset=64).tolist() == [1.0, 2.0, 3.0]
    assert np.frombuffer(raw, dtype="<u4", count=2, offset=88).tolist() == [3, 0]
    assert np.frombuffer(raw, dtype="<f8", count=1, offset=96).tolist() == [0.0]
    assert np.frombuffer(raw, dtype="<u4", count=1, offset=104).tolist() == [0]
    assert raw[108:128] == b"\0" * 20
    assert np.array_equal(np.frombuffer(raw, dtype="<f8", offset=128), np.arange(24.0))
    assert len(raw) == 128 + 24 * 8


@pytest.mark.parametrize("mode", [0, 1])
def test_writer_matches_reference_writer_byte_for_byte(oracle_built, tmp_path, mode):
    if not oracle_built.ref_available():
        pytest.skip("needs oracle/_ref (the reference's own writer)")
    ours, theirs = str(tmp_path / "ours.grid"), str(tmp_path / "ref.grid")
    v = _values()
    oracle_built.ref_save_file(mode, theirs, COUNTS, SPACING, ORIGIN, v, grid_type="ljr", inv_power=2.0, inv_power_mode=2)
    gf.write_grid_file(ours, COUNTS, SPACING, ORIGIN, v, grid_type=2 if mode == 0 else 0, inv_power=2.0, inv_power_mode=2,
                       with_trailer=(mode == 1))       # GridData::saveToFile writes grid type 0 (GridData.cpp:219-221)
    assert open(ours, "rb").read() == open(theirs, "rb").read()


@pytest.mark.parametrize("with_trailer", [False, True])
def test_reference_reader_reads_our_files(oracle_built, tmp_path, with_trailer):
    if not oracle_built.ref_available():
        pytest.skip("needs oracle/_ref (the reference's own reader)")
    path = str(tmp_path / "g.grid")
    v = _values()
    gf.write_grid_file(path, COUNTS, SPACING, ORIGIN, v, inv_power=3.0, inv_power_mode=2, with_trailer=with_trailer)
    counts, spacing, origin, vals, ip, mode = oracle_built.ref_load_file(path, v.size)
    assert counts == COUNTS and spacing == SPACING and origin == ORIGIN and (ip, mode) == (3.0, 2)
    assert np.array_equal(vals, v)


def test_our_reader_reads_reference_files(oracle_built, tmp_path):
    if not oracle_built.ref_available():
        pytest.skip("needs oracle/_ref")
    v = _values()
    for mode in (0, 1):
        path = str(tmp_path / f"ref{mode}.grid")
        oracle_built.ref_save_file(mode, path, COUNTS, SPACING, ORIGIN, v, grid_type="charge")
        h, w = gf.read_grid_file(path)
        assert h["counts"] == COUNTS and h["spacing"] == SPACING and h["origin"] == ORIGIN and np.array_equal(v, w)
        assert h["grid_type"] == (1 if mode == 0 else 0)


def test_bad_files_raise(tmp_path):
    bad = tmp_path / "bad.grid"
    bad.write_bytes(b"NOTAGRID" + b"\0" * 200)
    with pytest.raises(gf.GridForceB200Error, match="bad magic"):
        gf.read_grid_file(str(bad))
    path = str(tmp_path / "v2.grid")
    gf.write_grid_file(path, (2, 2, 2), (1, 1, 1), (0, 0, 0), np.zeros(8))
    raw = bytearray(open(path, "rb").read())
    raw[8] = 2                                         # version 2
    open(path, "wb").write(raw)
    with pytest.raises(gf.GridForceB200Error, match="Only V3"):
        gf.read_grid_file(path)
    raw[8] = 3
    open(path, "wb").write(raw[:150])                  # truncated data
    with pytest.raises(gf.GridForceB200Error, match="ends before"):
        gf.read_grid_file(path)
    with pytest.raises(gf.GridForceB200Error, match="Cannot open"):
        gf.read_grid_file(str(tmp_path / "missing.grid"))


@pytest.mark.gpu
def test_grid_from_file_on_device(gpu_device, tmp_path):
    """gfb_grid_create_from_file streams the file through pinned staging; evaluation equals the in-memory upload."""
    rng = np.random.default_rng(5)
    counts, sp, og = (70, 41, 33), (0.0125, 0.02, 0.03), (0.1, -0.2, 0.3)
    v = rng.normal(size=counts)
    path = str(tmp_path / "big.grid")
    gf.write_grid_file(path, counts, sp, og, v, with_trailer=True)
    g_file = gf.Grid.from_file(gpu_device, path, gf.PRECISION_DOUBLE)
    g_mem = gf.Grid(gpu_device, counts, sp, og, v, gf.PRECISION_DOUBLE)
    assert g_file.counts == counts and g_file.device_bytes == g_mem.device_bytes
    pos = np.array(og) + rng.uniform(0, 1, size=(500, 3)) * (np.array(sp) * (np.array(counts) - 1))
    sc = rng.normal(size=(1, 500))
    out = []
    for g in (g_file, g_mem):
        k = gf.Kernel(gpu_device, [g], sc)
        out.append(k.execute_host(pos))
        k.close()
    assert out[0][0][0] == out[1][0][0] and np.array_equal(out[0][1], out[1][1])
    g_file.close()
    g_mem.close()
