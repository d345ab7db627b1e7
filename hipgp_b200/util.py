"""Host-side index maps of the block-diagonal family (drop-in for `define_block_chunks`, ziggy/misc/util.py:79-126).

The maps are pure integer bookkeeping built once per model on the host (bit-exact with the reference by construction:
same chunking of each axis, same row-major flattening, blocks enumerated axis 0 outermost); the arithmetic that uses
them runs in `hipgp_block_lam` / `hipgp_block_diag_multiply`, which read k_n THROUGH the map instead of permuting it.
"""
import numpy as np
import torch


def define_block_chunks(xgrids, chunk_sizes):
    """neighbouring chunks of `prod(chunk_sizes)` grid points.  Returns (blk_idx (num_blocks, block_size) int64,
    to_blocks, from_blocks) like the reference; to_blocks / from_blocks are plain index gathers kept for API parity."""
    ndim = len(xgrids)
    assert ndim == len(chunk_sizes), "xgrids ndim = {}, chunk_sizes ndim = {}".format(ndim, len(chunk_sizes))
    assert ndim in (2, 3), "only 2d or 3d inputs"
    lens = [len(x) for x in xgrids]
    for d, (n, c) in enumerate(zip(lens, chunk_sizes)):
        assert n % c == 0, "xgrid-{}={} not divis by chunk_size={}".format(d, n, c)
    strides = [int(np.prod(lens[d + 1:])) for d in range(ndim)]
    # flat index of every grid point, reshaped to (n0/c0, c0, n1/c1, c1, ...) and the chunk axes moved to the back
    flat = np.arange(int(np.prod(lens)), dtype=np.int64).reshape(lens)
    shape = []
    for n, c in zip(lens, chunk_sizes):
        shape += [n // c, c]
    t = flat.reshape(shape)
    order = [2 * d for d in range(ndim)] + [2 * d + 1 for d in range(ndim)]
    blk = np.ascontiguousarray(t.transpose(order)).reshape(-1, int(np.prod(chunk_sizes)))
    assert strides[-1] == 1
    blk_idx = torch.from_numpy(blk)

    def to_blocks(m):
        return m[..., blk_idx.to(m.device)]

    flat_rev = torch.argsort(blk_idx.flatten())

    def from_blocks(block_m):
        return block_m.flatten(start_dim=1)[..., flat_rev.to(block_m.device)]

    return blk_idx, to_blocks, from_blocks
