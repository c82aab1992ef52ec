"""CPU: the C++17 host layer without a GPU -- (1) libcrcnn_b200_host.so loads and exports every entry point crcnn_b200/host.py binds;
(2) the split ShardedNetwork uses (crcnn_b200.hpp: shard_range) is the contiguous balanced one the gather-order tests assume
(nets.shard_range), including more ranks than outputs; (3) a network that reaches the reference's re-encryption point without a policy
refuses to run BEFORE it touches the device (no silent divergence from CrCNN/src/network.cpp:30)."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r'''
#include <cstdio>
#include "crcnn_b200/cpp/cnn_builder.hpp"
using namespace crcnn_b200;
struct Dummy : Layer {
    DeviceTensor forward_dev(DeviceTensor in) override { return in; }
    void printLayerStructure() override {}
    void savePlaintextParameters(std::ostream *) override {}
    void loadPlaintextParameters(std::istream *) override {}
};
int main() {
    for (int total : {1, 2, 7, 10, 20, 50, 500, 63}) for (int world : {1, 2, 3, 4, 8, 16}) {
        int covered = 0;
        for (int r = 0; r < world; r++) { int f, c; ShardedNetwork::shard_range(total, world, r, &f, &c); std::printf("%d %d %d %d %d\n", total, world, r, f, c); covered += c; }
        if (covered != total) return 3;
    }
    Network net;
    for (int i = 0; i < 7; i++) net.getLayers().push_back(std::shared_ptr<Layer>(new Dummy()));
    bool refused = false;
    try { net.forward(ciphertext3D(1, ciphertext2D(1, std::vector<Ciphertext>(1)))); } catch (const std::logic_error &e) { refused = std::string(e.what()).find("re-encrypts before layer 6") != std::string::npos; }
    std::printf("refused %d needs %d\n", (int)refused, (int)net.needs_reencryption());
    net.getLayers().resize(6);
    std::printf("six_layers_need %d\n", (int)net.needs_reencryption());
    int z, x, y, o; CnnBuilder::topologyShape("PlainModel", &z, &x, &y, &o);
    std::printf("shape %d %d %d %d\n", z, x, y, o);
    return 0;
}
'''


def test_host_library_exports():
    from crcnn_b200 import host
    h = host.load()
    for name in ("crcnn_host_init", "crcnn_host_set_evk", "crcnn_host_build", "crcnn_host_shape", "crcnn_host_layer_name",
                 "crcnn_host_forward_range", "crcnn_host_resident_begin", "crcnn_host_resident_run", "crcnn_host_resident_end",
                 "crcnn_host_serve", "crcnn_host_serve_times", "crcnn_host_set_fusion", "crcnn_host_ctx", "crcnn_host_shutdown", "crcnn_host_last_error"):
        assert hasattr(h, name), name
    # without a network the calls fail with a message instead of crashing
    v = [ctypes.c_int() for _ in range(5)]
    assert h.crcnn_host_shape(*[ctypes.byref(x) for x in v]) != 0 and b"no network" in h.crcnn_host_last_error()


def test_shard_split_and_reencryption_policy(tmp_path):
    from crcnn_b200 import nets
    src, exe = tmp_path / "t.cpp", tmp_path / "t"
    src.write_text(PROG)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + ROOT, "-o", str(exe), str(src), "-L" + os.path.join(ROOT, "crcnn_b200"),
                           "-lcrcnn_b200", "-Wl,-rpath," + os.path.join(ROOT, "crcnn_b200")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-500:] + out.stderr[-500:]
    lines = out.stdout.strip().splitlines()
    for ln in lines:
        f = ln.split()
        if len(f) == 5 and f[0].isdigit():
            total, world, r, first, count = map(int, f)
            assert nets.shard_range(total, world, r) == (first, count), ln
    assert "refused 1 needs 1" in lines and "six_layers_need 0" in lines and "shape 1 32 32 10" in lines


def test_reference_harness_is_not_stale():
    """oracle/_ref/dropin_seal_test (the reference's own classes next to the drop-in ones, run by tests/test_gpu_cpp_dropin.py on the GPU box,
    where /root/reference does not exist) must have been rebuilt after the last change to the headers it compiles: `make -q` on the recipe."""
    import subprocess
    if not os.path.isdir("/root/reference") or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "dropin_seal_test")):
        pytest.skip("no reference tree / harness not built here")
    rc = subprocess.run(["make", "-q", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref"], capture_output=True).returncode
    assert rc == 0, "oracle/_ref is older than its sources: run `python -c 'import __graft_entry__ as g; g.build()'`"
