#!/usr/bin/env python
"""One-off converter for the reference's *data* files (no code is copied).

The image has no HDF5/netCDF4 reader, and /root/reference does not exist on the GPU box, so the two
kinds of data the hot path needs are converted once, here, into plain .npz files that travel with the repo:

  * pyspeedy/data/example_bc.nc  (HDF5/NetCDF-4; 12 single-chunk shuffle+deflate float32 datasets)
        -> pyspeedy_b200/data/example_bc.npz     (boundary conditions read by Speedy.set_bc)
  * pyspeedy/tests/fixtures/1982-01-0{2,4}_0000.nc (NetCDF-3 classic)
        -> tests/golden/fixture_1982-01-0{2,4}.npz (the reference's own golden outputs, test_speedy.py:27-50)

The HDF5 file is decoded without an HDF5 library: every dataset is one deflate stream, so we scan the file
for zlib streams, inflate them, undo the HDF5 shuffle filter (element size 4) and keep the streams whose
inflated size matches a (96,48) or (96,48,12) float32 array.  The stream order in the file was established
in SURVEY.md section 7.1(6); it is re-verified below by physical-range checks on each field.
"""
import os
import sys
import zlib

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ORDER = ["orog", "lsm", "vegh", "alb", "vegl", "stl", "snowd", "swl1", "swl2", "swl3", "icec", "sst"]
FILL = np.float32(9.96921e36)


def unshuffle(raw, esize=4):
    n = len(raw) // esize
    a = np.frombuffer(raw, dtype=np.uint8).reshape(esize, n)
    return a.T.copy().tobytes()


def find_streams(buf):
    out = []
    pos = 0
    n = len(buf)
    while pos < n - 2:
        if buf[pos] == 0x78 and buf[pos + 1] in (0x01, 0x5E, 0x9C, 0xDA):
            d = zlib.decompressobj()
            try:
                data = d.decompress(buf[pos:])
            except zlib.error:
                pos += 1
                continue
            if d.eof and len(data) in (96 * 48 * 4, 96 * 48 * 12 * 4):
                used = n - pos - len(d.unused_data)
                out.append((pos, used, data))
                pos += used
                continue
        pos += 1
    return out


def convert_bc():
    src = os.path.join(REF, "pyspeedy/data/example_bc.nc")
    buf = open(src, "rb").read()
    streams = find_streams(buf)
    assert len(streams) == 12, len(streams)
    fields = {}
    for name, (pos, used, data) in zip(ORDER, streams):
        arr = np.frombuffer(unshuffle(data), dtype="<f4")
        if arr.size == 96 * 48:
            arr = arr.reshape(96, 48)
        else:
            arr = arr.reshape(96, 48, 12)
        arr = arr.copy()
        # xarray decodes _FillValue to NaN (pyspeedy/speedy.py:277 uses xr.load_dataset)
        arr[arr == FILL] = np.nan
        fields[name] = arr
        print(f"{name:6s} offset={pos:7d} clen={used:6d} shape={arr.shape} "
              f"min={np.nanmin(arr):.4g} max={np.nanmax(arr):.4g} nan={int(np.isnan(arr).sum())}")
    # plausibility checks that pin the name <-> stream mapping
    assert -50 < np.nanmin(fields["orog"]) < 0 and 4000 < np.nanmax(fields["orog"]) < 7000
    for k in ("lsm", "vegh", "vegl", "icec"):
        assert np.nanmin(fields[k]) >= -1e-6 and np.nanmax(fields[k]) <= 1 + 1e-6, k
    assert 0.0 < np.nanmin(fields["alb"]) and np.nanmax(fields["alb"]) < 1
    assert 200 < np.nanmin(fields["stl"]) and np.nanmax(fields["stl"]) < 330
    assert 230 < np.nanmin(fields["sst"]) and np.nanmax(fields["sst"]) < 310
    assert np.nanmin(fields["snowd"]) >= 0 and np.nanmax(fields["snowd"]) > 100
    for k in ("swl1", "swl2", "swl3"):
        assert np.nanmin(fields[k]) >= 0 and np.nanmax(fields[k]) < 1
    # lsm mean over globe ~0.3 ; land fields are missing over sea and vice versa
    assert 0.2 < fields["lsm"].mean() < 0.4
    assert np.isnan(fields["sst"]).sum() < np.isnan(fields["stl"]).sum()
    dst = os.path.join(ROOT, "pyspeedy_b200/data/example_bc.npz")
    np.savez_compressed(dst, **fields)
    print("wrote", dst, os.path.getsize(dst))


def convert_fixtures():
    import scipy.io

    for day in ("02", "04"):
        src = os.path.join(REF, f"pyspeedy/tests/fixtures/1982-01-{day}_0000.nc")
        f = scipy.io.netcdf_file(src, "r", mmap=False)
        out = {k: np.array(v.data, dtype=v.data.dtype.newbyteorder("=")) for k, v in f.variables.items()}
        dst = os.path.join(ROOT, f"tests/golden/fixture_1982-01-{day}.npz")
        np.savez_compressed(dst, **out)
        print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference not mounted; the committed .npz files are the outputs of this script")
    convert_bc()
    convert_fixtures()
