"""Writes float32 arrays as a version-0-superblock HDF5 file with the object kinds h5py's default settings produce
(symbol-table root group, group B-tree, symbol node, local heap, version-1 object headers, contiguous layout) -- the
fixture generator for the tests of crcnn_b200/cpp/h5lite.hpp on machines where the reference's own .h5 files are absent.
Laid out from the HDF5 File Format Specification; the reader is ALSO checked against the real h5py-written files of the
reference where those exist (tests/test_h5lite.py)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + bytes((-len(b)) % 8)


def _msg(mtype, body):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _dataset_header(shape, address, nbytes):
    space = struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)
    # IEEE little-endian float32: class 1 version 1; bit field: byte order 0, padding 0, mantissa normalisation 2 (implied), sign bit 31
    dtype = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", address, nbytes)
    msgs = _msg(0x0001, space) + _msg(0x0003, dtype) + _msg(0x0008, layout)
    return struct.pack("<BBHII4x", 1, 0, 3, 1, len(msgs)) + msgs


def write_h5(path, arrays):
    names = sorted(arrays)
    assert len(names) <= 32, "one symbol node only"
    # local heap data segment: "" at offset 0, then the names
    heap = bytearray(8)
    name_off = {}
    for nme in names:
        name_off[nme] = len(heap)
        heap += _pad8(nme.encode() + b"\0")
    heap = bytes(heap) + bytes(8)
    SUPER, ROOT_HDR = 0, 96
    root_msgs_len = 8 + 16
    at = ROOT_HDR + 16 + root_msgs_len
    btree_at = at
    at += 8 + 16 + (2 * 32 + 1) * 8            # node header + 2K+1 keys / 2K children for K = 16
    heap_hdr_at = at
    at += 32
    heap_data_at = at
    at += len(heap)
    snod_at = at
    at += 8 + 32 * 40
    hdr_at, data_at = {}, {}
    for nme in names:
        hdr_at[nme] = at
        at += len(_dataset_header(arrays[nme].shape, 0, 0))
    for nme in names:
        at = (at + 7) & ~7
        data_at[nme] = at
        at += arrays[nme].size * 4
    eof = at
    out = bytearray(eof)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 16, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, ROOT_HDR, 1, 0) + struct.pack("<QQ", btree_at, heap_hdr_at)
    out[SUPER:SUPER + len(sb)] = sb
    root = struct.pack("<BBHII4x", 1, 0, 1, 1, root_msgs_len) + _msg(0x0011, struct.pack("<QQ", btree_at, heap_hdr_at))
    out[ROOT_HDR:ROOT_HDR + len(root)] = root
    tree = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF)
    tree += struct.pack("<QQQ", 0, snod_at, name_off[names[-1]] if names else 0)
    out[btree_at:btree_at + len(tree)] = tree
    hh = b"HEAP" + struct.pack("<B3x", 0) + struct.pack("<QQQ", len(heap), len(heap) - 8, heap_data_at)
    out[heap_hdr_at:heap_hdr_at + len(hh)] = hh
    out[heap_data_at:heap_data_at + len(heap)] = heap
    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for nme in names:
        sn += struct.pack("<QQII16x", name_off[nme], hdr_at[nme], 0, 0)
    out[snod_at:snod_at + len(sn)] = sn
    for nme in names:
        a = np.ascontiguousarray(arrays[nme], dtype="<f4")
        h = _dataset_header(a.shape, data_at[nme], a.size * 4)
        out[hdr_at[nme]:hdr_at[nme] + len(h)] = h
        out[data_at[nme]:data_at[nme] + a.size * 4] = a.tobytes()
    open(path, "wb").write(bytes(out))
