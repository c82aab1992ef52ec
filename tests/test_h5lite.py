"""CPU: the HDF5 reader behind CnnBuilder::getPretrained (crcnn_b200/cpp/h5lite.hpp).

* against the reference's own h5py-written weight files where /root/reference exists (this container): every dataset equal,
  bit for bit, to weights/*.npz (exported from the .pth state dicts, tools/export_weights.py);
* against files written by tests/h5write.py from the committed weights (runs anywhere)."""
import os
import subprocess

import numpy as np
import pytest

from h5write import write_h5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/PlainModel"
MODELS = ["PlainModel", "ApproxPlainModel", "PlainModelWoPad", "PlainModelTiny"]


@pytest.fixture(scope="module")
def dump_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("h5") / "h5lite_dump")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "h5lite_dump.cpp")])
    return exe


def check(dump_exe, path, want, tmp_path):
    res = subprocess.run([dump_exe, path, str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    listed = {}
    for line in res.stdout.splitlines():
        f = line.split()
        rank = int(f[1])
        listed[f[0]] = (tuple(int(v) for v in f[2:2 + rank]), int(f[2 + rank]), int(f[3 + rank]))
    for name, a in want.items():
        assert listed[name] == (a.shape, 1, 4), name
        got = np.fromfile(os.path.join(str(tmp_path), name + ".f32"), dtype=np.float32)
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(a, dtype=np.float32).ravel().view(np.uint32)), name


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's .h5 files are only in the dev container")
@pytest.mark.parametrize("model", MODELS)
def test_reads_reference_files(dump_exe, model, tmp_path):
    check(dump_exe, os.path.join(REF, model + ".h5"), dict(np.load(os.path.join(ROOT, "weights", model + ".npz"))), tmp_path)


def test_reads_written_fixture(dump_exe, tmp_path):
    rng = np.random.default_rng(3)
    want = {"a.weight": rng.standard_normal((3, 1, 5, 5)).astype(np.float32), "a.bias": rng.standard_normal(3).astype(np.float32),
            "z": np.float32([[1.5, -2.25]]), "scalar_like": np.float32([7])}
    path = str(tmp_path / "w.h5")
    write_h5(path, want)
    check(dump_exe, path, want, tmp_path)


def test_rejects_garbage(dump_exe, tmp_path):
    path = str(tmp_path / "bad.h5")
    open(path, "wb").write(b"not an hdf5 file at all" * 10)
    res = subprocess.run([dump_exe, path, str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 1 and "not an HDF5 file" in res.stderr
