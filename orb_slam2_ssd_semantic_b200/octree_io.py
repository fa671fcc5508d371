"""OctoMap `.ot` (ColorOcTree) export / import of the GPU leaf map.

Reference: MapDrawer::SaveOctoMap -> m_octree->write(name) (perfect/src/MapDrawer.cc:1103-1111) after InsertScan's
prune() (:1024).  File layout (octomap AbstractOcTree::write + OcTreeBase::writeData; pinned by the reference's own
artefact `octomap.ot`: 390 133 nodes x 8 bytes): text header, then the tree in pre-order, each node =
float32 log-odds + 3 x u8 colour + one byte whose bit i says child i exists.  Child index at depth d (root = 0) of a
16-bit OcTreeKey: bit (15-d) of x | y<<1 | z<<2 (octomap computeChildIdx).  Inner node value = max of its children
(updateOccupancyChildren); a node whose 8 children are identical childless leaves is pruned into one leaf carrying the
child's data (pruneNode).  Inner-node colours are history dependent in octomap (a node keeps the colour it got while
it was a pruned leaf after it is expanded again; the artefact shows exactly that: only some inner nodes are
non-white), so they cannot be derived from the final leaves: the exporter writes white (255,255,255), octomap's
constructor value.  tests/test_oracle_cpu.py round-trips the reference's octomap.ot: node count, order, every
log-odds value, every child mask and every leaf colour are reproduced bit for bit.
"""
from __future__ import annotations

import io
from typing import Tuple

import numpy as np

HEADER = ("# Octomap OcTree file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n"
          "id ColorOcTree\nsize %d\nres %s\ndata\n")
NODE_DT = np.dtype([("v", "<f4"), ("rgb", "u1", 3), ("child", "u1")])
DEPTH = 16


def read_ot(path_or_bytes):
    """-> (res, nodes[NODE_DT] in file order, header_text)."""
    b = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    i = b.index(b"data\n") + 5
    header = b[:i].decode("ascii", "replace")
    size = int([ln for ln in header.splitlines() if ln.startswith("size")][0].split()[1])
    res_txt = [ln for ln in header.splitlines() if ln.startswith("res")][0].split()[1]
    nodes = np.frombuffer(b[i:i + size * NODE_DT.itemsize], dtype=NODE_DT)
    return float(res_txt), nodes, header


def leaves_from_nodes(nodes: np.ndarray):
    """Walk the pre-order node array -> arrays (keys[n,3] u16 of the leaf's min corner, depth[n], value[n], rgb[n,3])
    for every childless node, plus is_consistent (inner value == max child value everywhere)."""
    keys, depths, vals, cols = [], [], [], []
    consistent = True
    pos = 0
    stack = [(0, 0, 0, 0)]     # (kx, ky, kz, depth) of the node about to be read, pushed in reverse child order
    maxchild_stack = []        # (node_index, remaining_children, running max)
    v_all, child_all, rgb_all = nodes["v"], nodes["child"], nodes["rgb"]
    n = len(nodes)
    while stack:
        kx, ky, kz, d = stack.pop()
        if pos >= n:
            raise ValueError("truncated .ot")
        me = pos
        pos += 1
        cm = int(child_all[me])
        # account this node as a child of the innermost open parent
        if cm == 0:
            keys.append((kx, ky, kz)); depths.append(d); vals.append(v_all[me]); cols.append(rgb_all[me])
            val_for_parent = v_all[me]
            while maxchild_stack:
                pi, rem, mx = maxchild_stack[-1]
                mx = max(mx, val_for_parent)
                rem -= 1
                if rem > 0:
                    maxchild_stack[-1] = (pi, rem, mx)
                    break
                maxchild_stack.pop()
                if v_all[pi] != mx:
                    consistent = False
                val_for_parent = v_all[pi]
        else:
            cnt = bin(cm).count("1")
            maxchild_stack.append((me, cnt, -np.inf))
            bit = 1 << (DEPTH - 1 - d)
            for c in range(7, -1, -1):
                if cm & (1 << c):
                    stack.append((kx | (bit if c & 1 else 0), ky | (bit if c & 2 else 0), kz | (bit if c & 4 else 0), d + 1))
    return (np.array(keys, np.uint16).reshape(-1, 3), np.array(depths, np.int32), np.array(vals, np.float32),
            np.array(cols, np.uint8).reshape(-1, 3), consistent, pos)


def expand_to_max_depth(keys, depths, vals, cols):
    """Pruned leaves (depth < 16) -> their depth-16 cells (what the GPU hash map stores)."""
    out_k, out_v, out_c = [keys[depths == DEPTH]], [vals[depths == DEPTH]], [cols[depths == DEPTH]]
    for i in np.nonzero(depths < DEPTH)[0]:
        side = 1 << (DEPTH - int(depths[i]))
        g = np.arange(side, dtype=np.uint32)
        xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
        k = np.stack([xx.ravel() + keys[i, 0], yy.ravel() + keys[i, 1], zz.ravel() + keys[i, 2]], 1).astype(np.uint16)
        out_k.append(k)
        out_v.append(np.full(len(k), vals[i], np.float32))
        out_c.append(np.repeat(cols[i][None, :], len(k), 0))
    return np.concatenate(out_k), np.concatenate(out_v), np.concatenate(out_c)


def _morton_order(keys: np.ndarray) -> np.ndarray:
    """Sort key so that octomap's pre-order (child 0..7 at every depth, bit order x | y<<1 | z<<2, MSB first) is
    ascending: interleave the 16 bits of (x, y, z) with x as the least significant of each triple."""
    k = keys.astype(np.uint64)
    code = np.zeros(len(keys), np.uint64)
    for b in range(DEPTH):
        code |= ((k[:, 0] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        code |= ((k[:, 1] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + 1)
        code |= ((k[:, 2] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + 2)
    return code


def build_ot(keys: np.ndarray, logodds: np.ndarray, rgb: np.ndarray, res: float, res_text: str | None = None) -> bytes:
    """Depth-16 leaves -> bytes of a pruned ColorOcTree `.ot` file (same canonical form octomap writes after prune())."""
    keys = np.ascontiguousarray(keys, np.uint16).reshape(-1, 3)
    logodds = np.ascontiguousarray(logodds, np.float32)
    rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
    if len(keys) == 0:
        return (HEADER % (0, res_text or repr(res))).encode()
    code = _morton_order(keys)
    order = np.argsort(code, kind="stable")
    code, logodds, rgb = code[order], logodds[order], rgb[order]
    # level arrays bottom-up: nodes[d] = (prefix code >> 3*(16-d)), value, rgb, childmask, is_leaf
    lvl_code, lvl_val, lvl_rgb = code, logodds, rgb
    lvl_mask = np.zeros(len(code), np.uint8)
    levels = [None] * (DEPTH + 1)
    levels[DEPTH] = (lvl_code, lvl_val, lvl_rgb, lvl_mask)
    for d in range(DEPTH, 0, -1):
        c_code, c_val, c_rgb, c_mask = levels[d]
        parent = c_code >> np.uint64(3)
        uniq, start, counts = np.unique(parent, return_index=True, return_counts=True)
        child_idx = (c_code & np.uint64(7)).astype(np.uint8)
        p_mask = np.zeros(len(uniq), np.uint8)
        np.bitwise_or.at(p_mask, np.repeat(np.arange(len(uniq)), counts), (np.uint8(1) << child_idx))
        p_val = np.maximum.reduceat(c_val, start)
        p_rgb = np.full((len(uniq), 3), 255, np.uint8)                      # inner colours: constructor white
        # prune: 8 childless children with identical value and colour -> the parent becomes that leaf
        first_val, first_rgb = c_val[start], c_rgb[start]
        same_val = np.minimum.reduceat(c_val, start) == p_val
        leafless = np.maximum.reduceat(c_mask, start) == 0
        rgb_key = c_rgb[:, 0].astype(np.int64) << 16 | c_rgb[:, 1].astype(np.int64) << 8 | c_rgb[:, 2].astype(np.int64)
        same_rgb = np.minimum.reduceat(rgb_key, start) == np.maximum.reduceat(rgb_key, start)
        prune = (counts == 8) & same_val & leafless & same_rgb
        p_mask[prune] = 0
        p_rgb[prune] = first_rgb[prune]
        p_val = np.where(prune, first_val, p_val).astype(np.float32)
        # drop the children of pruned parents from level d
        keep = ~np.repeat(prune, counts)
        levels[d] = (c_code[keep], c_val[keep], c_rgb[keep], c_mask[keep])
        levels[d - 1] = (uniq, p_val, p_rgb, p_mask)
    # pre-order emission: sort all nodes by (prefix code aligned to 48 bits, depth)
    all_code, all_d, all_val, all_rgb, all_mask = [], [], [], [], []
    for d in range(DEPTH + 1):
        c_code, c_val, c_rgb, c_mask = levels[d]
        all_code.append(c_code << np.uint64(3 * (DEPTH - d)))
        all_d.append(np.full(len(c_code), d, np.int64))
        all_val.append(c_val); all_rgb.append(c_rgb); all_mask.append(c_mask)
    all_code = np.concatenate(all_code); all_d = np.concatenate(all_d)
    order = np.lexsort((all_d, all_code))
    out = np.zeros(len(order), NODE_DT)
    out["v"] = np.concatenate(all_val)[order]
    out["rgb"] = np.concatenate(all_rgb)[order]
    out["child"] = np.concatenate(all_mask)[order]
    buf = io.BytesIO()
    buf.write((HEADER % (len(out), res_text or repr(res))).encode())
    buf.write(out.tobytes())
    return buf.getvalue()


def save_octomap(pcm, path: str) -> Tuple[int, int]:
    """PointCloudMapping (GPU leaf map) -> `.ot` file readable by octovis.  Returns (leaves, nodes written)."""
    keys, lo, rgb = pcm.export_leaves()
    data = build_ot(keys, lo, rgb, pcm.resolution)
    with open(path, "wb") as f:
        f.write(data)
    _, nodes, _ = read_ot(data)
    return len(keys), len(nodes)
