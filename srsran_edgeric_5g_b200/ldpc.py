"""Host-side LDPC helpers: TS 38.212 segmentation arithmetic used to describe codeblocks to the GPU.

Mirrors the reference's inline helpers (include/srsran/phy/upper/channel_coding/ldpc/ldpc.h:128-228) and the rx half of
ldpc_segmenter_impl (lib/phy/upper/channel_coding/ldpc/ldpc_segmenter_impl.cpp:58-68, :254-331): cheap integer
arithmetic that stays on the host, exactly as SURVEY 8a R18 prescribes.
"""
from dataclasses import dataclass

LIFTING_SIZES = (2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52,
                 56, 60, 64, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352,
                 384)
MAX_CODEBLOCK_SIZE = 66 * 384
MAX_BITS_CRC16 = 3824
BG1, BG2 = 1, 2


def compute_tb_crc_size(tbs_bits):
    return 16 if tbs_bits <= MAX_BITS_CRC16 else 24


def compute_nof_codeblocks(tbs_bits, base_graph):
    b = tbs_bits + compute_tb_crc_size(tbs_bits)
    max_seg = 8448 if base_graph == BG1 else 3840
    return 1 if b <= max_seg else -(-b // (max_seg - 24))


def compute_lifting_size(tbs_bits, base_graph, nof_segments):
    b = tbs_bits + compute_tb_crc_size(tbs_bits)
    ref = 22
    if base_graph == BG2:
        ref = 10 if b > 640 else 9 if b > 560 else 8 if b > 192 else 6
    b_out = b + (24 * nof_segments if nof_segments > 1 else 0)
    for z in LIFTING_SIZES:
        if z * ref * nof_segments >= b_out:
            return z
    raise ValueError("transport block too large")


def compute_codeblock_size(base_graph, lifting_size):
    return (22 if base_graph == BG1 else 10) * lifting_size


def compute_full_codeblock_size(base_graph, codeblock_size):
    return codeblock_size * (3 if base_graph == BG1 else 5)


def compute_N_ref(tbs_lbrm_bytes, nof_codeblocks):
    return min(tbs_lbrm_bytes * 8 * 3 // (2 * nof_codeblocks), MAX_CODEBLOCK_SIZE)


@dataclass
class CodeblockMetadata:
    """codeblock_metadata (include/srsran/phy/upper/codeblock_metadata.h:42-80), flattened."""
    base_graph: int
    lifting_size: int
    rv: int
    mod: int  # bits per symbol
    Nref: int
    cw_length: int
    full_length: int
    rm_length: int
    nof_filler_bits: int
    cw_offset: int
    nof_crc_bits: int


def segment_rx(tbs_bits, base_graph, rv, mod, Nref, nof_layers, nof_cw_llrs):
    """ldpc_segmenter_rx::segment: per-codeblock metadata of a received codeword."""
    assert nof_cw_llrs % mod == 0 and (nof_cw_llrs // mod) % nof_layers == 0
    C = compute_nof_codeblocks(tbs_bits, base_graph)
    b_in = tbs_bits + compute_tb_crc_size(tbs_bits)
    b_out = b_in + (24 * C if C > 1 else 0)
    Z = compute_lifting_size(tbs_bits, base_graph, C)
    K = compute_codeblock_size(base_graph, Z)
    crc_bits = 24 if C > 1 else 0
    max_info = -(-b_out // C) - crc_bits
    sym_per_layer = (nof_cw_llrs // mod) // nof_layers
    n_short = C - (sym_per_layer % C)
    out, offset = [], 0
    for i in range(C):
        per_cb = sym_per_layer // C if i < n_short else -(-sym_per_layer // C)
        rm = per_cb * nof_layers * mod
        out.append(
            CodeblockMetadata(base_graph, Z, rv, mod, Nref, nof_cw_llrs, compute_full_codeblock_size(base_graph, K), rm,
                              K - (max_info + crc_bits), offset,
                              compute_tb_crc_size(tbs_bits) if C == 1 else 24))
        offset += rm
    assert offset == nof_cw_llrs
    return out


def segment_tx(ctx, transport_block, base_graph):
    """ldpc_segmenter_tx::segment (lib/phy/upper/channel_coding/ldpc/ldpc_segmenter_tx_impl.cpp; TS 38.212 5.2.2, 7.2.3):
    transport-block CRC attachment, segmentation, codeblock CRC attachment and filler bits. Returns one array of K bits
    per codeblock (filler bits as zeros: the rate matcher skips them by position). CRCs come from the device's CRC
    calculator (ctx.crc = pdc_crc)."""
    import numpy as np
    from . import capi
    tb = np.ascontiguousarray(transport_block, np.uint8)
    tbs_bits = tb.size * 8
    tcrc = compute_tb_crc_size(tbs_bits)
    chk = ctx.crc(capi.CRC16 if tcrc == 16 else capi.CRC24A, tb, tbs_bits)
    bits = np.concatenate([np.unpackbits(tb), np.array([(chk >> (tcrc - 1 - i)) & 1 for i in range(tcrc)], np.uint8)])
    C = compute_nof_codeblocks(tbs_bits, base_graph)
    Z = compute_lifting_size(tbs_bits, base_graph, C)
    K = compute_codeblock_size(base_graph, Z)
    cb_crc = 24 if C > 1 else 0
    b_out = bits.size + cb_crc * C
    info = -(-b_out // C) - cb_crc  # information bits per codeblock; the last one is zero padded
    segments, off = [], 0
    for _ in range(C):
        seg = np.zeros(K, np.uint8)
        take = min(info, bits.size - off)
        seg[:take] = bits[off:off + take]
        off += take
        if cb_crc:
            c = ctx.crc(capi.CRC24B, np.packbits(seg[:info]), info)
            seg[info:info + 24] = [(c >> (23 - i)) & 1 for i in range(24)]
        segments.append(seg)
    return segments
