"""Multi-GPU plumbing for the block path: blocks are independent (io/CompressedStream.go:896-898), so they are sharded
round-robin over ranks with no collective on the compute path. torch.distributed (NCCL over NVLink on GPUs, gloo in
the CPU tests) is only used to gather the variable-length encoded blocks on rank 0, where they are committed in block
order exactly like Writer.processBlock does (5-bit / lw-bit prefixes, io/CompressedStream.go:951-976).
"""
import numpy as np

BS_TYPE = 0x4B414E5A
BS_VERSION = 6


def shard_blocks(nblocks, world, rank):
    """Block ids handled by `rank`: round-robin (BASELINE.json north_star)."""
    return list(range(rank, nblocks, world))


class BitAssembler:
    """MSB-first bit string builder (DefaultOutputBitStream semantics) used for the ordered commit on rank 0."""

    def __init__(self):
        self.parts = []
        self.nbits = 0

    def write_bits(self, value, count):
        bits = np.array([(value >> (count - 1 - i)) & 1 for i in range(count)], np.uint8)
        self.parts.append(bits)
        self.nbits += count

    def write_array(self, data, nbits):
        if nbits:
            a = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else data
            self.parts.append(np.unpackbits(a[: (nbits + 7) // 8])[:nbits])
            self.nbits += nbits

    def to_bytes(self):
        allbits = np.concatenate(self.parts) if self.parts else np.zeros(0, np.uint8)
        return np.packbits(allbits)


def header_checksum(ck_size, etype, t48, block_size, sz_mask, input_size):
    """Writer.writeHeader :498-512."""
    M = 0xFFFFFFFF
    H = 0x1E35A7BD
    ck = (H * ((0x01030507 * BS_VERSION) & M)) & M
    ck ^= (H * (~ck_size & M)) & M
    ck ^= (H * (~etype & M)) & M
    ck ^= (H * ((~t48 >> 32) & M)) & M
    ck ^= (H * (~t48 & M)) & M
    ck ^= (H * (~block_size & M)) & M
    if sz_mask:
        ck ^= (H * ((~input_size >> 32) & M)) & M
        ck ^= (H * (~input_size & M)) & M
    return ((ck >> 23) ^ (ck >> 3)) & 0xFFFFFF


def assemble_stream(blocks, transform48, entropy, block_size, input_size=0, checksum_bits=0):
    """blocks: ordered list of (bytes, nbits) block-local bit strings -> complete stream (header, blocks, end marker)."""
    out = BitAssembler()
    ck_size = {0: 0, 32: 1, 64: 2}[checksum_bits]
    out.write_bits(BS_TYPE, 32)
    out.write_bits(BS_VERSION, 4)
    out.write_bits(ck_size, 2)
    out.write_bits(entropy, 5)
    out.write_bits(transform48, 48)
    out.write_bits(block_size >> 4, 28)
    if input_size <= 0 or input_size >= (1 << 48):
        sz_mask = 0
    elif input_size >= (1 << 32):
        sz_mask = 3
    elif input_size >= (1 << 16):
        sz_mask = 2
    else:
        sz_mask = 1
    out.write_bits(sz_mask, 2)
    if sz_mask:
        out.write_bits(input_size, 16 * sz_mask)
    out.write_bits(0, 15)
    out.write_bits(header_checksum(ck_size, entropy, transform48, block_size, sz_mask, input_size if sz_mask else 0), 24)
    for data, nbits in blocks:
        lw = 3 if nbits < 8 else (nbits >> 3).bit_length() - 1 + 4
        out.write_bits(lw - 3, 5)
        out.write_bits(nbits, lw)
        out.write_array(data, nbits)
    out.write_bits(0, 5)
    out.write_bits(0, 3)
    return out.to_bytes()


def gather_blocks(local, nblocks, dist, device="cpu"):
    """local: list of (block_id, bytes array, nbits) produced by this rank. Returns on rank 0 the ordered list
    [(bytes, nbits)] of all nblocks blocks (None elsewhere). Two collectives: all_gather of the bit lengths (fixed
    size) and one all_gather of the padded payloads (gather-v)."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    bits = torch.zeros(nblocks, dtype=torch.int64, device=device)
    for b, _, nb in local:
        bits[b] = nb
    dist.all_reduce(bits)  # every block has exactly one owner
    sizes = (bits + 7) // 8
    per_rank = [int(sum(int(sizes[b]) for b in shard_blocks(nblocks, world, r))) for r in range(world)]
    cap = max(per_rank + [1])
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    o = 0
    for b, data, nb in sorted(local):
        n = (nb + 7) // 8
        buf[o:o + n] = torch.as_tensor(np.ascontiguousarray(data[:n]), device=device)
        o += n
    bufs = [torch.zeros(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != 0:
        return None
    out = [None] * nblocks
    for r in range(world):
        o = 0
        flat = bufs[r].cpu().numpy()
        for b in shard_blocks(nblocks, world, r):
            n = int(sizes[b])
            out[b] = (flat[o:o + n].copy(), int(bits[b]))
            o += n
    return out
