"""bigfile mesh reader (reference: nbodykit BigFileMesh at scripts/measure/measure_bs_slow.py:226).

The fixture below is written byte by byte from the library's documented block layout
(header / attr-v2 / hex-named data files), independently of bskit_b200.bigfile.save_mesh."""
import os

import numpy as np
import pytest

from bskit_b200 import bigfile as bf


def _write_fixture(root, arr, box, nfile, dtype):
    d = os.path.join(root, "Field")
    os.makedirs(d)
    flat = arr.astype(dtype).reshape(-1)
    cut = [len(flat) * i // nfile for i in range(nfile + 1)]
    head = "DTYPE: %s\nNMEMB: 1\nNFILE: %d\n" % (np.dtype(dtype).str, nfile)
    for i in range(nfile):
        raw = flat[cut[i]:cut[i + 1]].tobytes()
        with open(os.path.join(d, "%06X" % i), "wb") as f:
            f.write(raw)
        s = sum(raw)
        s = (s & 0xFFFF) + ((s & 0xFFFFFFFF) >> 16)
        s = (s & 0xFFFF) + (s >> 16)
        head += "%06X: %d : %d : 0\n" % (i, cut[i + 1] - cut[i], s)
    with open(os.path.join(d, "header"), "w") as f:
        f.write(head)
    shp = np.array(arr.shape, dtype="<i8")
    bx = np.array([box] * 3, dtype="<f8")
    with open(os.path.join(d, "attr-v2"), "w") as f:
        f.write("ndarray.shape <i8 3 %s #HUMANE [ %d %d %d ]\n" % (shp.tobytes().hex(), *arr.shape))
        f.write("BoxSize <f8 3 %s #HUMANE [ %g %g %g ]\n" % (bx.tobytes().hex(), box, box, box))
        f.write("Nmesh <i8 3 %s #HUMANE [ %d %d %d ]\n" % (shp.tobytes().hex(), *arr.shape))
        f.write("painted |S1 4 %s #HUMANE [ true ]\n" % b"true".hex())


@pytest.mark.parametrize("nfile,dtype", [(1, "<f4"), (3, "<f4"), (2, ">f8")])
def test_reads_handwritten_fixture(tmp_path, nfile, dtype):
    rng = np.random.default_rng(3)
    arr = rng.standard_normal((8, 8, 8))
    _write_fixture(str(tmp_path / "m.bigfile"), arr, 250.0, nfile, dtype)
    m = bf.BigFileMesh(str(tmp_path / "m.bigfile"), "Field", verify=True)
    want = arr.astype(dtype)
    assert np.array_equal(np.asarray(m.compute()), want)
    assert np.array_equal(m.attrs["BoxSize"], [250.0] * 3) and np.array_equal(m.attrs["Nmesh"], [8, 8, 8])
    assert m.attrs["painted"] == "true"
    assert np.array_equal(m.slab(2, 5), want[2:5])          # an x-slab, as one rank of a multi-GPU job reads it
    assert np.array_equal(np.asarray(m.array[4:8]), want[4:8])


def test_roundtrip_and_errors(tmp_path):
    arr = np.random.default_rng(1).standard_normal((6, 6, 6)).astype(np.float32)
    p = bf.save_mesh(str(tmp_path / "w.bigfile"), arr, 100.0, nfile=4, attrs={"seed": np.array([7])})
    m = bf.BigFileMesh(p, verify=True)
    assert np.array_equal(np.asarray(m.compute()), arr) and int(m.attrs["seed"][0]) == 7
    with pytest.raises(FileNotFoundError):
        bf.BigFileMesh(p, "Nope")
    with open(os.path.join(p, "Field", "000001"), "r+b") as f:      # corrupt one byte
        f.seek(3)
        b = f.read(1)
        f.seek(3)
        f.write(bytes([b[0] ^ 0x5A]))
    with pytest.raises(ValueError, match="checksum"):
        bf.BigFileMesh(p, verify=True)


def test_cast_source_accepts_bigfile(tmp_path):
    from bskit_b200.mesh import cast_source
    arr = np.zeros((4, 4, 4), dtype=np.float32)
    p = bf.save_mesh(str(tmp_path / "c.bigfile"), arr, [10.0, 10.0, 10.0])
    m = cast_source(bf.BigFileMesh(p))
    assert int(m.attrs["Nmesh"][0]) == 4 and m.array.shape == (4, 4, 4)
    with pytest.raises(ValueError):
        cast_source(bf.BigFileMesh(p), BoxSize=11.0)
