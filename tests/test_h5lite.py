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


def test_malformed_files_are_rejected_not_trusted(tmp_path):
    """Sizes read from the file are never trusted: element size 64, absurd length size, dimensions that wrap around,
    truncation and 400 random single-byte corruptions of a valid file must end in a clean error or a clean read -- under
    AddressSanitizer / UBSan, so an out-of-bounds access fails the test instead of passing silently."""
    exe = str(tmp_path / "h5lite_dump_asan")
    src = os.path.join(ROOT, "tests", "cpp", "h5lite_dump.cpp")
    san = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe, src],
                         capture_output=True, text=True)
    if san.returncode != 0:   # no sanitizer runtime in this image: plain build, crashes still show as negative return codes
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, src])
    rng = np.random.default_rng(11)
    want = {"w": rng.standard_normal((4, 6)).astype(np.float32), "b": rng.standard_normal(4).astype(np.float32)}
    good = str(tmp_path / "good.h5")
    write_h5(good, want)
    raw = bytearray(open(good, "rb").read())
    out = tmp_path / "out"
    out.mkdir()

    def run(data):
        path = str(tmp_path / "case.h5")
        open(path, "wb").write(bytes(data))
        res = subprocess.run([exe, path, str(out)], capture_output=True, text=True, timeout=60)
        assert res.returncode in (0, 1, 3), "reader crashed (rc %d): %s" % (res.returncode, res.stderr[-800:])
        return res

    assert run(raw).returncode == 0
    bad = bytearray(raw); bad[14] = 200                      # size of lengths
    assert run(bad).returncode == 1
    bad = bytearray(raw); bad[13] = 3                        # size of offsets
    assert run(bad).returncode == 1
    for cut in (50, 97, 300, len(raw) // 2, len(raw) - 5):   # truncation
        run(raw[:cut])
    # datatype message: class/size field.  Find the float32 datatype message (class 1, size 4) and blow the size up
    hits = [i for i in range(len(raw) - 8) if raw[i] == 0x11 and raw[i + 4:i + 8] == bytes([4, 0, 0, 0])]
    assert hits
    for i in hits:
        for size in (64, 0, 3, 0x7fffffff):
            bad = bytearray(raw); bad[i + 4:i + 8] = int(size).to_bytes(4, "little")
            assert run(bad).returncode == 1
    # dataspace dimensions that overflow 64 bits when multiplied
    dims = [i for i in range(len(raw) - 16) if raw[i:i + 8] == (4).to_bytes(8, "little") and raw[i + 8:i + 16] == (6).to_bytes(8, "little")]
    for i in dims:
        bad = bytearray(raw); bad[i:i + 8] = (1 << 62).to_bytes(8, "little"); bad[i + 8:i + 16] = (1 << 62).to_bytes(8, "little")
        assert run(bad).returncode == 1
    for k in range(400):                                     # random single-byte corruption
        bad = bytearray(raw)
        pos = int(rng.integers(0, len(raw)))
        bad[pos] = int(rng.integers(0, 256))
        run(bad)
