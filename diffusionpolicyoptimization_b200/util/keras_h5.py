"""`.weights.h5` checkpoints in the reference's layout, without h5py / TensorFlow.

The reference saves and loads Keras-3 weight files: `self.model.save_weights("state_{itr}.weights.h5")` for the fine-tuned
`PPODiffusion` (agent/finetune/train_agent.py:127-142), `self.model.network.save_weights(...)` / `ema_state_*` for the
pre-trained `DiffusionMLP` (agent/pretrain/train_agent.py:150-162), `load_weights(network_path)` in
model/diffusion/diffusion_vpg.py:92-97.  Neither h5py nor Keras is installed in the build image, so this module carries

  * a small pure-Python HDF5 subset (`H5Reader`, `write_h5`): superblock version 0, old-style groups (symbol tables: v1 B-tree
    of group nodes + local heap), object headers version 1 with continuation blocks, contiguous / compact / single-chunk
    little-endian float datasets - what h5py writes with its default `libver` for files like these, and what it reads back;
  * the Keras-3 variable paths of the reference's model classes (`keras_paths_*`).

LAYOUT ASSUMPTION (cannot be validated here - no Keras): Keras 3 `saving_lib._save_state` walks the attributes of a
saveable in sorted order (`_walk_saveable`), stores a layer's own variables as datasets "0", "1", ... under
"<path>/vars", names the items of a list attribute by the snake-cased class name (+ "_<n>" from the second occurrence)
and the layers of a `Sequential` under "layers/".  For the reference's classes that gives, with Dense = (kernel [in,out],
bias):

    DiffusionMLP   time_embedding/layers/dense/vars/{0,1}, time_embedding/layers/dense_1/vars/{0,1},
                   mlp_mean/input_layer/vars/{0,1},
                   mlp_mean/residual_blocks/two_layer_pre_activation_res_net_linear/{l1,l2}/vars/{0,1},
                   mlp_mean/output_layer/vars/{0,1}
    CriticObs      Q1/input_layer/..., Q1/residual_blocks/two_layer_pre_activation_res_net_linear/{l1,l2}/..., Q1/output_layer/...
    PPODiffusion   actor/<DiffusionMLP>, actor_ft/<DiffusionMLP>, critic/<CriticObs>   (`network` is the same object as `actor`)

The flat variable ORDER used everywhere else in this package is the creation order documented in include/dppo_b200.h.
`load_keras_weights_h5` looks every variable up by path; if a file uses other group names it falls back to matching the
file's datasets (in file order) against the expected shapes, and says so.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


# ------------------------------------------------------------------------------------------------ Keras paths
def keras_paths_diffusion_mlp(prefix: str = "") -> List[str]:
    """Dataset paths of a residual-style DiffusionMLP in the flat variable order of this package
    (time Dense32 W, b, time Dense16 W, b, input W, b, block.l1 W, b, block.l2 W, b, output W, b)."""
    p = prefix.rstrip("/") + "/" if prefix else ""
    blk = p + "mlp_mean/residual_blocks/two_layer_pre_activation_res_net_linear/"
    layers = [p + "time_embedding/layers/dense/", p + "time_embedding/layers/dense_1/", p + "mlp_mean/input_layer/",
              blk + "l1/", blk + "l2/", p + "mlp_mean/output_layer/"]
    return [layer + "vars/" + str(i) for layer in layers for i in (0, 1)]


def keras_paths_critic_obs(prefix: str = "") -> List[str]:
    p = prefix.rstrip("/") + "/" if prefix else ""
    blk = p + "Q1/residual_blocks/two_layer_pre_activation_res_net_linear/"
    layers = [p + "Q1/input_layer/", blk + "l1/", blk + "l2/", p + "Q1/output_layer/"]
    return [layer + "vars/" + str(i) for layer in layers for i in (0, 1)]


# ------------------------------------------------------------------------------------------------ writer
class _Buf:
    def __init__(self):
        self.b = bytearray()

    def tell(self):
        return len(self.b)

    def align(self, n=8):
        while len(self.b) % n:
            self.b.append(0)

    def write(self, data: bytes):
        off = len(self.b)
        self.b += data
        return off

    def patch(self, off: int, data: bytes):
        self.b[off:off + len(data)] = data


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    pad = (-len(body)) % 8
    return struct.pack("<HHB3x", mtype, len(body) + pad, flags) + body + b"\0" * pad


def _object_header(msgs: Sequence[bytes]) -> bytes:
    body = b"".join(msgs)
    # version 1, reserved, #messages, object reference count, header size; 4 bytes of padding align the messages to 8
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


def _f32_datatype(dt: np.dtype) -> bytes:
    if dt == np.float32:
        return struct.pack("<B3BI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt == np.float64:
        return struct.pack("<B3BI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    raise TypeError(f"unsupported dtype {dt}")


def write_h5(path: str, datasets: Dict[str, np.ndarray]) -> None:
    """Write float32 / float64 arrays at slash-separated paths (groups are created as needed)."""
    tree: dict = {}
    for p, arr in datasets.items():
        node = tree
        parts = [x for x in p.split("/") if x]
        for g in parts[:-1]:
            node = node.setdefault(g, {})
            if not isinstance(node, dict):
                raise ValueError(f"{p}: a dataset is in the way")
        node[parts[-1]] = np.ascontiguousarray(arr)
    buf = _Buf()
    buf.write(b"\0" * 96)                                        # superblock (56 bytes) + root symbol table entry (40 bytes)

    def write_dataset(arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype=arr.dtype.newbyteorder("<"))
        buf.align()
        data_addr = buf.write(arr.tobytes()) if arr.size else UNDEF
        dims = arr.shape
        space = struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)               # version 3, contiguous
        fill = struct.pack("<BBBB", 2, 2, 2, 0)                                  # version 2, late allocation, fill if set, undefined
        hdr = _object_header([_msg(0x0001, space), _msg(0x0003, _f32_datatype(arr.dtype), 1), _msg(0x0005, fill, 1), _msg(0x0008, layout)])
        buf.align()
        return buf.write(hdr)

    def write_group(node: dict) -> Tuple[int, int, int]:
        """-> (object header address, B-tree address, heap address)"""
        names = sorted(node)                                     # symbol tables are ordered by name
        child_addr = {}
        for n in names:
            child_addr[n] = write_group(node[n])[0] if isinstance(node[n], dict) else write_dataset(node[n])
        # local heap: offset 0 holds the empty string, then the link names, each padded to 8 bytes
        heap_data = bytearray(b"\0" * 8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap_data)
            raw = n.encode() + b"\0"
            heap_data += raw + b"\0" * ((-len(raw)) % 8)
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)                   # one free block: next = 1 (end of list), size 16
        buf.align()
        heap_seg = buf.write(bytes(heap_data))
        buf.align()
        heap_addr = buf.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_seg))
        # symbol table nodes of at most 8 entries (2 x leaf K = 4), one B-tree node over them (at most 32 children)
        snods = []
        for i in range(0, max(len(names), 1), 8):
            chunk = names[i:i + 8]
            ent = b"".join(struct.pack("<QQII16x", name_off[n], child_addr[n], 0, 0) for n in chunk)
            ent += b"\0" * (40 * (8 - len(chunk)))
            buf.align()
            snods.append((buf.write(b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)) + ent), chunk))
        if len(snods) > 32:
            raise ValueError("too many links in one group for this writer")
        keys = [0] + [name_off[ch[-1]] if ch else 0 for _, ch in snods]
        body = b"".join(struct.pack("<QQ", keys[i], snods[i][0]) for i in range(len(snods))) + struct.pack("<Q", keys[-1])
        node_bytes = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + body
        node_bytes += b"\0" * (24 + (2 * 16) * 8 + (2 * 16 + 1) * 8 - len(node_bytes))       # full-size node (internal K = 16)
        buf.align()
        btree_addr = buf.write(node_bytes)
        buf.align()
        hdr_addr = buf.write(_object_header([_msg(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]))
        return hdr_addr, btree_addr, heap_addr

    root_hdr, root_btree, root_heap = write_group(tree)
    buf.align()
    eof = buf.tell()
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    root_entry = struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
    buf.patch(0, sb + root_entry)
    with open(path, "wb") as f:
        f.write(bytes(buf.b))


# ------------------------------------------------------------------------------------------------ reader
class H5Reader:
    """Datasets of an HDF5 file written with old-style groups (superblock 0 / 1), in file (name) order."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.d = f.read()
        base = self.d.find(SIG)
        if base != 0:
            raise ValueError("not an HDF5 file (or a user block precedes the superblock)")
        ver = self.d[8]
        if ver not in (0, 1):
            raise ValueError(f"HDF5 superblock version {ver} (new-style groups) is not supported by this reader: re-save with h5py's default libver")
        if self.d[13] != 8 or self.d[14] != 8:
            raise ValueError("only 8-byte offsets / lengths are supported")
        off = 24 + (4 if ver == 1 else 0)
        off += 32                                                # base, free-space, end-of-file, driver-info addresses
        _, root_hdr, cache, _ = struct.unpack_from("<QQII", self.d, off)
        self.datasets: Dict[str, np.ndarray] = {}
        self._walk(root_hdr, "")

    # ---- object headers
    def _messages(self, addr: int):
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.d, addr)
        if ver != 1:
            raise ValueError(f"object header version {ver} is not supported by this reader")
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.d, p)
                body = self.d[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                              # continuation block
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr, clen))
                out.append((mtype, body))
        return out

    def _walk(self, hdr_addr: int, prefix: str):
        msgs = self._messages(hdr_addr)
        types = {t for t, _ in msgs}
        if 0x0011 in types:
            body = next(b for t, b in msgs if t == 0x0011)
            btree, heap = struct.unpack_from("<QQ", body, 0)
            for name, child in self._group_entries(btree, heap):
                self._walk(child, prefix + "/" + name if prefix else name)
        elif 0x0008 in types:
            self.datasets[prefix] = self._dataset(msgs)
        # anything else (committed datatypes, new-style groups) is ignored

    def _group_entries(self, btree: int, heap: int):
        if self.d[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        _, _, seg = struct.unpack_from("<QQQ", self.d, heap + 8)
        out = []

        def node(addr):
            if self.d[addr:addr + 4] == b"SNOD":
                n = struct.unpack_from("<H", self.d, addr + 6)[0]
                for i in range(n):
                    name_off, obj = struct.unpack_from("<QQ", self.d, addr + 8 + 40 * i)
                    end = self.d.index(b"\0", seg + name_off)
                    out.append((self.d[seg + name_off:end].decode(), obj))
                return
            if self.d[addr:addr + 4] != b"TREE":
                raise ValueError("bad B-tree signature")
            _ntype, _level, used = struct.unpack_from("<BBH", self.d, addr + 4)
            p = addr + 24
            for i in range(used):
                child = struct.unpack_from("<Q", self.d, p + 8 + 16 * i)[0]
                node(child)
        node(btree)
        return out

    def _dataset(self, msgs) -> np.ndarray:
        space = next(b for t, b in msgs if t == 0x0001)
        dtype = next(b for t, b in msgs if t == 0x0003)
        layout = next(b for t, b in msgs if t == 0x0008)
        sver, rank, sflags = struct.unpack_from("<BBB", space, 0)
        doff = 8 if sver == 1 else 4
        dims = struct.unpack_from("<" + "Q" * rank, space, doff) if rank else ()
        cls = dtype[0] & 0x0F
        size = struct.unpack_from("<I", dtype, 4)[0]
        big = dtype[1] & 1
        if cls == 1:
            np_dt = np.dtype(("<" if not big else ">") + "f" + str(size))
        elif cls == 0:
            signed = (dtype[1] >> 3) & 1
            np_dt = np.dtype(("<" if not big else ">") + ("i" if signed else "u") + str(size))
        else:
            raise ValueError(f"datatype class {cls} is not supported by this reader")
        n = int(np.prod(dims)) if rank else 1
        lver = layout[0]
        if lver == 3:
            lclass = layout[1]
            if lclass == 1:                                      # contiguous
                addr, _sz = struct.unpack_from("<QQ", layout, 2)
                raw = self.d[addr:addr + n * size] if addr != UNDEF else b"\0" * (n * size)
            elif lclass == 0:                                    # compact
                sz = struct.unpack_from("<H", layout, 2)[0]
                raw = layout[4:4 + sz]
            elif lclass == 2:                                    # chunked: one chunk covering the dataset, no filters
                crank = layout[2]
                btree = struct.unpack_from("<Q", layout, 3)[0]
                raw = self._single_chunk(btree, crank, n * size)
            else:
                raise ValueError("unknown data layout class")
        else:
            raise ValueError(f"data layout message version {lver} is not supported by this reader")
        return np.frombuffer(raw, dtype=np_dt, count=n).reshape(dims).astype(np_dt.newbyteorder("="))

    def _single_chunk(self, btree: int, crank: int, nbytes: int) -> bytes:
        if self.d[btree:btree + 4] != b"TREE":
            raise ValueError("bad chunk B-tree")
        ntype, level, used = struct.unpack_from("<BBH", self.d, btree + 4)
        if ntype != 1 or level != 0 or used != 1:
            raise ValueError("chunked datasets with more than one chunk are not supported by this reader")
        p = btree + 24
        csize, fmask = struct.unpack_from("<II", self.d, p)
        if fmask != 0:
            pass
        addr = struct.unpack_from("<Q", self.d, p + 8 + 8 * crank)[0]
        if csize < nbytes:
            raise ValueError("filtered (compressed) chunks are not supported by this reader")
        return self.d[addr:addr + nbytes]


# ------------------------------------------------------------------------------------------------ package-level helpers
def save_keras_weights_h5(path: str, weights: Sequence[np.ndarray], var_paths: Sequence[str]) -> None:
    if len(weights) != len(var_paths):
        raise ValueError("one path per variable is needed")
    write_h5(path, {p: np.asarray(w, np.float32) for p, w in zip(var_paths, weights)})


def load_keras_weights_h5(path: str, var_paths: Sequence[str], shapes: Sequence[Tuple[int, ...]]) -> List[np.ndarray]:
    r = H5Reader(path)
    if all(p in r.datasets for p in var_paths):
        out = [np.asarray(r.datasets[p], np.float32) for p in var_paths]
    else:
        # other group names (another Keras version): match the file's float datasets against the expected shapes
        pool = [(k, v) for k, v in r.datasets.items() if v.dtype.kind == "f"]
        out, used = [], set()
        for shp in shapes:
            hit = next((i for i, (_, v) in enumerate(pool) if i not in used and tuple(v.shape) == tuple(shp)), None)
            if hit is None:
                raise KeyError(f"{path}: no dataset of shape {tuple(shp)} left; the file holds {[(k, v.shape) for k, v in pool]}")
            used.add(hit)
            out.append(np.asarray(pool[hit][1], np.float32))
        import logging
        logging.getLogger(__name__).warning("%s: variable paths differ from the assumed Keras-3 layout; matched %d datasets by shape in file order",
                                            path, len(out))
    for w, shp in zip(out, shapes):
        if tuple(w.shape) != tuple(shp):
            raise ValueError(f"{path}: variable has shape {w.shape}, the network expects {tuple(shp)}")
    return out
