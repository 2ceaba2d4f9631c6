#!/usr/bin/env python3
"""Randomised parity soak on a GPU box: for a number of minutes, random rate-dematcher geometries (new data and
retransmissions, all modulation orders, limited buffers, fillers, non-finite soft bits) and random decoder batches (both
base graphs, every lifting size, 0.2-1.1 laps of soft bits, early stop on / off, 1-8 iterations) against the oracle,
bit for bit: HARQ entries, decoded bits, CRC flags, iteration counts; and randomised HARQ histories (populations of
entries living through slots of new transmissions and retransmissions, tests/vectors.py: harq_sequence_rounds). Prints one JSON line; exit code 1 on a mismatch.

    python tools/soak.py [minutes] [seed]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    from oracle import pyoracle as po
    from srsran_edgeric_5g_b200 import capi
    from tests.vectors import LIFTING_SIZES, harq_sequence_rounds, make_cb_batch
    orc = po.Oracle()
    ctx = capi.Context(device=0, max_cbs=256, max_llrs=256 * 3 * 25344, harq_entries=256, max_tbs=4, max_tb_bytes=1 << 16)
    # HARQ histories through the queued batch path: large populations run one dematcher CTA per codeblock
    ctx_seq = capi.Context(device=0, max_cbs=1400, max_llrs=64 << 20, harq_entries=1400, max_tbs=64, max_tb_bytes=8 << 20)
    from srsran_edgeric_5g_b200.pusch_decoder import rx_buffer_pool
    from tests.test_gpu_pusch_decoder import tb_generations
    pool_seq = rx_buffer_pool(ctx_seq, first_entry=0, nof_entries=1400)  # the same entries as the codeblock histories
    rng = np.random.default_rng(seed)
    t_end = time.time() + 60.0 * minutes
    n_dm = n_dec = n_seq = n_tb = 0
    while time.time() < t_end:
        # ---- transport-block level: generations of UEs with other transport block sizes on the same pool keys
        n_ue, max_tb = [(30, 5000), (12, 60000), (40, 20000)][int(rng.integers(0, 3))]
        try:
            n_tb += tb_generations(ctx_seq, orc, rng, pool_seq, n_gen=2, n_ue=n_ue, max_tb_bytes_bg1=max_tb)
        except AssertionError as err:
            print(json.dumps({"mismatch": "transport block generations", "case": repr(err.args)[:2000], "seed": seed}))
            return 1
        # ---- HARQ histories: new transmissions and retransmissions of random geometry on entries that live on
        n_ent, max_Z = [(1300, 64), (700, 128), (60, 384)][int(rng.integers(0, 3))]
        prev_dbg = locals().get("dbg", {})
        dbg = {}
        try:
            n_seq += harq_sequence_rounds(ctx_seq, orc, rng, n_ent=n_ent, rounds=4, max_Z=max_Z, debug=dbg)
        except AssertionError as err:
            info = {"mismatch": "harq sequence", "case": repr(err.args)[:2000], "seed": seed, "n_ent": n_ent}
            if dbg:
                e = dbg["ent"]
                gb, wb = np.unpackbits(dbg["got"])[:e["K"]], np.unpackbits(dbg["want"])[:e["K"]]
                diff = np.nonzero(gb != wb)[0]
                nz = np.nonzero(e["buf"])[0]
                info.update(result=repr(dbg["res"]), differing_bits=int(diff.size), first=diff[:16].tolist(), last=diff[-4:].tolist(),
                            last_nonzero=int(nz[-1]), before_nonzero=np.nonzero(dbg["before"])[0][:20].tolist(),
                            before_count=int(np.count_nonzero(dbg["before"])), after_tail=e["buf"][1560:1640].tolist(), neighbours=[repr(c) for c in dbg["cbs"][max(0, dbg["index"] - 2):dbg["index"] + 3]])
                info["previous_population"] = [repr(h[dbg["index"]]) for h in prev_dbg.get("history", []) if dbg["index"] < h.size]
                info["this_population"] = [repr(h[dbg["index"]]) for h in dbg.get("history", [])]
            print(json.dumps(info))
            return 1
        # ---- rate dematcher, single calls
        for _ in range(50):
            bg = int(rng.integers(1, 3))
            Z = int(rng.choice(LIFTING_SIZES))
            kb = 22 if bg == 1 else 10
            N = (66 if bg == 1 else 50) * Z
            Ksys = (kb - 2) * Z
            qm = int(rng.choice([1, 2, 4, 6, 8]))
            F = int(rng.integers(0, min(Ksys - 1, 2 * Z))) if rng.random() < 0.6 else 0
            if rng.random() < 0.5:
                F -= F % 4
            nref = int(rng.integers(Ksys + 2 * Z, N + 50)) if rng.random() < 0.4 else 0
            E = int(rng.integers(1, max(2, 3 * N // qm))) * qm
            if rng.random() < 0.5:
                E -= E % (4 * qm)
                E = max(E, 4 * qm)
            rv = int(rng.integers(0, 4)) if rng.random() < 0.6 else 0
            buf0 = (rng.integers(-120, 121, N) if rng.random() < 0.7 else rng.integers(-128, 128, N)).astype(np.int8)
            llr = (rng.integers(-120, 121, E) if rng.random() < 0.8 else rng.integers(-128, 128, E)).astype(np.int8)
            new_data = bool(rng.integers(0, 2))
            a, b = buf0.copy(), buf0.copy()
            ctx.rate_dematch(a, llr, new_data, rv, qm, nref, F)
            orc.rate_dematch(b, llr, new_data, rv, qm, nref, F, 64)
            if not (a == b).all():
                print(json.dumps({"mismatch": "rate_dematch", "case": [bg, Z, qm, F, nref, E, rv, new_data], "seed": seed}))
                return 1
            n_dm += 1
        # ---- batches through dematcher + throughput decoder
        for _ in range(6):
            bg = int(rng.integers(1, 3))
            Z = int(rng.choice(LIFTING_SIZES[8:] if rng.random() < 0.7 else LIFTING_SIZES))
            n_short = 66 if bg == 1 else 50
            qm = int(rng.choice([2, 4, 6, 8]))
            E = int(rng.integers((24 if bg == 1 else 12), n_short + 8)) * Z
            E -= E % (4 * qm) if rng.random() < 0.7 else E % qm
            E = max(E, 4 * qm)
            F = int(rng.integers(0, Z))
            F -= F % 4 if rng.random() < 0.7 else 0
            rv = 0 if rng.random() < 0.7 else int(rng.integers(0, 4))
            n_cb = int(rng.integers(1, 24))
            b = make_cb_batch(orc, bg, Z, n_cb=n_cb, E=E, qm=qm, rv=rv, snr_db=float(rng.uniform(-3, 6)),
                              seed=int(rng.integers(1 << 30)), crc_kind=po.CRC24B if Z > 3 else po.CRC16, nof_filler=F)
            mi, es = int(rng.integers(1, 9)), bool(rng.integers(0, 2))
            out = b.run_gpu(ctx, mi, es, harq_init=0)
            ref = b.run_oracle(orc, mi, es)
            ok = ((out["crc_ok"] == ref["crc_ok"]).all() and (out["iters"] == ref["iters"]).all() and
                  (out["bits"] == ref["bits"]).all() and (out["harq"] == ref["harq"]).all())
            if not ok:
                print(json.dumps({"mismatch": "decode batch", "case": [bg, Z, qm, E, F, rv, n_cb, mi, es], "seed": seed}))
                return 1
            n_dec += n_cb
    print(json.dumps({"soak_ok": True, "minutes": minutes, "seed": seed, "rate_dematch_calls": n_dm,
                      "codeblocks_decoded": n_dec, "harq_sequence_codeblocks": n_seq, "transport_blocks_decoded": n_tb,
                      "canaries": min(ctx.debug_canaries_ok(), ctx_seq.debug_canaries_ok())}))
    ctx_seq.close()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
