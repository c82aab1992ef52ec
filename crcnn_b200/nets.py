"""Network topologies of the reference's builder and a device-resident predict path over the C ABI.

Topologies mirror CnnBuilder::buildNetwork (CrCNN/src/cnnBuilder.cpp:108-179: Approx, WoPad and
Tiny blocks); `PlainModel` is the weights file PlainModel.h5 run through the unchanged layer API on
a 32x32 zero-bordered input (SURVEY.md section 8(a), table of topologies, and Appendix B3).

Network.forward mirrors Network::forward (CrCNN/src/network.cpp:22-47) without the hard-coded
re-encryption before layer 6: re-encryption needs the SECRET key and therefore belongs to the
client; `forward(first, last)` exposes the segment API for it.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHTS_DIR = os.path.join(os.path.dirname(_HERE), "weights")

# layer tuples: (kind, name, args...) with the reference's constructor argument order
TOPOLOGIES = {
    # cnnBuilder.cpp:115-134
    "ApproxPlainModel": dict(input=(1, 28, 28), layers=[
        ("conv", "pool1_features.conv1", 28, 28, 1, 2, 2, 5, 5, 20),
        ("avgpool", "pool1", 12, 12, 20, 1, 1, 2, 2),
        ("bn", "pool1_features.norm1", 20, 11, 11),
        ("conv", "pool2_features.conv2", 11, 11, 20, 2, 2, 3, 3, 50),
        ("square", "act1", 50, 5, 5),
        ("avgpool", "pool2", 5, 5, 50, 1, 1, 2, 2),
        ("bn", "pool2_features.norm2", 50, 4, 4),
        ("fc", "classifier.fc3", 800, 500),
        ("fc", "classifier.fc4", 500, 10)]),
    # cnnBuilder.cpp:136-155
    "PlainModelWoPad": dict(input=(1, 28, 28), layers=[
        ("conv", "pool1_features.conv1", 28, 28, 1, 2, 2, 5, 5, 20),
        ("pool", "pool1", 12, 12, 20, 1, 1, 2, 2),
        ("bn", "pool1_features.norm1", 20, 11, 11),
        ("conv", "pool2_features.conv2", 11, 11, 20, 2, 2, 3, 3, 50),
        ("square", "act1", 50, 5, 5),
        ("pool", "pool2", 5, 5, 50, 1, 1, 2, 2),
        ("bn", "pool2_features.norm2", 50, 4, 4),
        ("fc", "classifier.fc3", 800, 500),
        ("fc", "classifier.fc4", 500, 10)]),
    # cnnBuilder.cpp:157-169
    "PlainModelTiny": dict(input=(1, 28, 28), layers=[
        ("conv", "pool1_features.conv1", 28, 28, 1, 1, 1, 5, 5, 32),
        ("avgpool", "pool1", 24, 24, 32, 2, 2, 2, 2),
        ("conv", "pool2_features.conv2", 12, 12, 32, 1, 1, 5, 5, 64),
        ("avgpool", "pool2", 8, 8, 64, 2, 2, 2, 2),
        ("fc", "classifier.fc3", 1024, 512),
        ("fc", "classifier.fc4", 512, 10)]),
    # PlainModel.h5: conv1 trained with padding=2 -> 32x32 zero-bordered input, fc3 is 1250 -> 500
    "PlainModel": dict(input=(1, 32, 32), layers=[
        ("conv", "pool1_features.conv1", 32, 32, 1, 2, 2, 5, 5, 20),
        ("avgpool", "pool1", 14, 14, 20, 1, 1, 2, 2),
        ("bn", "pool1_features.norm1", 20, 13, 13),
        ("conv", "pool2_features.conv2", 13, 13, 20, 2, 2, 3, 3, 50),
        ("square", "act1", 50, 6, 6),
        ("avgpool", "pool2", 6, 6, 50, 1, 1, 2, 2),
        ("bn", "pool2_features.norm2", 50, 5, 5),
        ("fc", "classifier.fc3", 1250, 500),
        ("fc", "classifier.fc4", 500, 10)]),
}


def load_weights(model):
    return dict(np.load(os.path.join(WEIGHTS_DIR, model + ".npz")))


def layer_terms(layer):
    """Weighted-sum terms (ciphertext x plaintext products) per image of one layer."""
    if layer[0] == "conv":
        _, _, xd, yd, zd, xs, ys, xf, yf, nf = layer
        return nf * ((xd - xf) // xs + 1) * ((yd - yf) // ys + 1) * zd * xf * yf
    if layer[0] == "fc":
        return layer[2] * layer[3]
    return 0


def layer_io_counts(layer):
    """(input ciphertexts, output ciphertexts) per image."""
    k = layer[0]
    if k == "conv":
        _, _, xd, yd, zd, xs, ys, xf, yf, nf = layer
        return zd * xd * yd, nf * ((xd - xf) // xs + 1) * ((yd - yf) // ys + 1)
    if k in ("pool", "avgpool"):
        _, _, xd, yd, zd, xs, ys, xf, yf = layer
        return zd * xd * yd, zd * ((xd - xf) // xs + 1) * ((yd - yf) // ys + 1)
    if k in ("bn", "square"):
        c = layer[2] * layer[3] * layer[4]
        return c, c
    if k == "fc":
        return layer[2], layer[3]
    raise ValueError(k)


class Network:
    """Encoded network resident on one GPU (weights as plaintext packs behind the C ABI)."""

    def __init__(self, eng, model, weights=None, evk=None, limit=None):
        """limit: optional dict {layer_name: max_outputs} to build a cropped network (used only by the
        bounded CPU-baseline sample; the GPU path always runs the full layers)."""
        self.eng, self.model = eng, model
        topo = TOPOLOGIES[model]
        self.input_shape = topo["input"]
        self.layers = list(topo["layers"])
        w = weights if weights is not None else load_weights(model)
        self.evk = evk
        self.packs = {}
        self.params = {}   # the float32 values every pack was encoded from (what a checker re-encodes)
        for layer in self.layers:
            kind, name = layer[0], layer[1]
            if kind in ("conv", "fc"):
                self.params[name] = (w[name + ".weight"].ravel().astype(np.float32), w[name + ".bias"].ravel().astype(np.float32))
                self.packs[name] = (eng.plain_encode(w[name + ".weight"].ravel()), eng.plain_encode(w[name + ".bias"].ravel()))
            elif kind == "bn":
                # CnnBuilder::buildBatchNormLayer, cnnBuilder.cpp:89-105 (float32 arithmetic as in the reference)
                var = w[name + ".running_var"].astype(np.float32)
                invstd = (1.0 / np.sqrt(var.astype(np.float64) + 0.00001)).astype(np.float32)  # double arithmetic, float result: cnnBuilder.cpp:101
                self.params[name] = (w[name + ".running_mean"].astype(np.float32), invstd)
                self.packs[name] = (eng.plain_encode(w[name + ".running_mean"]), eng.plain_encode(invstd))
            elif kind == "avgpool":
                xf, yf = layer[7], layer[8]
                self.packs[name] = (eng.plain_encode([1.0 / (xf * yf)]),)  # avgPoolingLayer.cpp:10-13

    def num_layers(self):
        return len(self.layers)

    def forward_layer(self, i, x, batch):
        eng, layer = self.eng, self.layers[i]
        kind, name = layer[0], layer[1]
        if kind == "conv":
            w, b = self.packs[name]
            return eng.conv(x, w, b, batch, *layer[2:])
        if kind == "fc":
            w, b = self.packs[name]
            return eng.fc(x, w, b, batch, layer[2], layer[3])
        if kind == "pool":
            return eng.pool(x, batch, *layer[2:])
        if kind == "avgpool":
            return eng.pool(x, batch, *layer[2:], scale=self.packs[name][0])
        if kind == "bn":
            m, v = self.packs[name]
            return eng.bn(x, batch, layer[2], layer[3], layer[4], m, v)
        if kind == "square":
            if self.evk is None:
                raise RuntimeError("square layer needs evaluation keys")
            return eng.square_layer(x, self.evk)
        raise ValueError(kind)

    def forward(self, x, batch=1, first=0, last=None, on_layer=None):
        """Layers [first, last) applied in order, activations staying on the device.
        on_layer(i, layer) is called after each layer has been enqueued (for event timing)."""
        last = len(self.layers) if last is None else last
        for i in range(first, last):
            y = self.forward_layer(i, x, batch)
            if i > first:
                x.free()  # intermediate activation
            x = y
            if on_layer is not None:
                on_layer(i, self.layers[i])
        return x


# ------------------------------------------------------------------------------------------------
# Output-neuron sharding: the reference splits a layer's output
# filters / rows across std::threads (convolutionalLayer.cpp:177-187, fullyConnectedLayer.cpp:148-158);
# across GPUs the same split applies, followed by an all-gather of the layer's output ciphertexts
# before the next layer that consumes every channel.
# ------------------------------------------------------------------------------------------------
def shard_range(total, world, rank):
    """Contiguous, balanced split of `total` output channels/rows: the first total % world ranks get one more."""
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_counts(total, world, per_channel):
    """Ciphertexts each rank contributes to the all-gather of a [channels][per_channel] activation."""
    return [shard_range(total, world, r)[1] * per_channel for r in range(world)]


class ShardedNetwork(Network):
    """One image (batch 1) with every conv / fc layer split by output channel / row across the ranks of a
    torch.distributed group; pool, batch-norm and square act on the local channels without any exchange;
    an all-gather of ciphertexts (NCCL over NVLink on GPUs) precedes every conv / fc layer."""

    def __init__(self, eng, model, dist, weights=None, evk=None):
        self.dist, self.rank, self.world = dist, dist.get_rank(), dist.get_world_size()
        super().__init__(eng, model, weights=weights, evk=evk)
        w = weights if weights is not None else load_weights(model)
        # per-rank slices of the per-channel parameters of batch-norm layers
        self.local_bn = {}
        channels = self.input_shape[0]
        for layer in self.layers:
            if layer[0] == "conv":
                channels = layer[9]
            elif layer[0] == "bn":
                k0, kc = shard_range(channels, self.world, self.rank)
                var = w[layer[1] + ".running_var"].astype(np.float32)
                invstd = (1.0 / np.sqrt(var.astype(np.float64) + 0.00001)).astype(np.float32)  # double arithmetic, float result: cnnBuilder.cpp:101
                self.local_bn[layer[1]] = (eng.plain_encode(w[layer[1] + ".running_mean"][k0:k0 + kc]),
                                           eng.plain_encode(invstd[k0:k0 + kc]))

    def all_gather(self, t, channels, per_channel):
        """Local [kc][per_channel] ciphertexts -> full [channels][per_channel] on every rank."""
        import torch
        eng = self.eng
        counts = gather_counts(channels, self.world, per_channel)
        words_per_ct = 2 * eng.K * eng.n
        full = eng.alloc(sum(counts), 2, ntt_form=eng.device_ptr(t)[1])
        # shards may differ by one channel: gather max-sized slots, then compact
        slot = max(counts) * words_per_ct
        src = _as_torch(eng.device_ptr(t)[0], counts[self.rank] * words_per_ct)
        dst = _as_torch(eng.device_ptr(full)[0], sum(counts) * words_per_ct)
        eng.sync()  # the engine's stream produced `t`; the collective runs on torch's stream
        padded = torch.empty(self.world * slot, dtype=torch.int64, device="cuda")
        mine = padded[self.rank * slot:self.rank * slot + counts[self.rank] * words_per_ct]
        mine.copy_(src)
        self.dist.all_gather_into_tensor(padded, padded[self.rank * slot:(self.rank + 1) * slot].clone())
        off = 0
        for r, c in enumerate(counts):
            dst[off:off + c * words_per_ct].copy_(padded[r * slot:r * slot + c * words_per_ct])
            off += c * words_per_ct
        torch.cuda.current_stream().synchronize()
        return full

    def forward(self, x, on_layer=None):
        eng = self.eng
        channels, sharded = self.input_shape[0], False  # `sharded`: x holds only this rank's channels
        for i, layer in enumerate(self.layers):
            kind, name = layer[0], layer[1]
            if kind in ("conv", "fc"):
                if sharded:
                    per = (layer[2] * layer[3]) if kind == "conv" else (layer[2] // channels)
                    y = self.all_gather(x, channels, per)
                    x.free()
                    x = y
                out_total = layer[9] if kind == "conv" else layer[3]
                k0, kc = shard_range(out_total, self.world, self.rank)
                wp, bp = self.packs[name]
                if kind == "conv":
                    y = eng.conv(x, wp, bp, 1, *layer[2:], shard=(k0, kc))
                else:
                    y = eng.fc(x, wp, bp, 1, layer[2], layer[3], shard=(k0, kc))
                channels, sharded = out_total, True
            else:
                k0, kc = shard_range(channels, self.world, self.rank) if sharded else (0, channels)
                if kind == "pool":
                    y = eng.pool(x, 1, layer[2], layer[3], kc, *layer[5:])
                elif kind == "avgpool":
                    y = eng.pool(x, 1, layer[2], layer[3], kc, *layer[5:], scale=self.packs[name][0])
                elif kind == "bn":
                    m, v = self.local_bn[name] if sharded else self.packs[name]
                    y = eng.bn(x, 1, kc, layer[3], layer[4], m, v)
                else:
                    y = eng.square_layer(x, self.evk)
            if i > 0:
                x.free()
            x = y
            if on_layer is not None:
                on_layer(i, layer)
        if sharded:  # final scores: gather the output rows
            y = self.all_gather(x, channels, 1)
            x.free()
            x = y
        return x


class _CudaView:
    def __init__(self, ptr, words):
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def _as_torch(ptr, words):
    import torch
    return torch.as_tensor(_CudaView(ptr, words), device="cuda")
