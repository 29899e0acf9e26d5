"""Per-layer work model of the SAO / O12 decoder as the plan executes it (CPU only; no GPU, no library).

For every launch of an inference plan in bf16 mode it lists the algorithmic FLOPs, the algorithmic HBM bytes (fp16
residual stream, bf16 operands), the bytes a tile pulls from L2 into shared memory (weights are re-streamed per
tile), the MMA cycles of a tile and the resulting L2->SM ingest in bytes per clock and SM at full MMA rate.  It is
the arithmetic behind DESIGN.md section 7 (which layers are bound by what, and what a CTA pair or an in-kernel operand
would buy); numbers measured on the GPU are in profiles/.

usage: python tools/layer_model.py [sao|o12] [clips]
"""
import sys

ARCH = {
    "sao": dict(latent=64, io=2, channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 8, 8], frames=216),
    "o12": dict(latent=512, io=1, channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 5, 8], frames=375),
}
MMA_FLOP_PER_CLK = 8192.0          # dense bf16 per SM
L2_TO_SM_CAP = 42.0                # B/clk/SM (B300_MICROARCH: ~6300 B/clk chip-wide TMA throughput / 148)


def decoder_layers(a, B):
    """(name, kind, Cin, Cout, K, stride, dilation, rows_out, extras) in launch order; fused ResidualUnits at C=128."""
    ch = [a["channels"] * m for m in a["c_mults"]]
    T = a["frames"]
    out = [("latent conv k7", "conv", a["latent"], ch[-1], 7, 1, 1, B * T, {})]
    c = ch[-1]
    for i in range(len(ch) - 1, -1, -1):
        s = a["strides"][i]
        co = ch[i - 1] if i > 0 else ch[0]
        k = 2 * s + s % 2
        T *= s
        out.append((f"convT s{s} {c}->{co}", "convT", c, co, k, s, 1, B * T, {"stream_out": True}))
        c = co
        for d in (1, 3, 9):
            last = d == 9
            if c == 128:
                out.append((f"fused RU d{d} C{c}", "ru", c, c, 7, 1, d, B * T, {"stream_out": not last}))
            else:
                out.append((f"RU d{d} k7 C{c}", "conv", c, c, 7, 1, d, B * T, {}))
                out.append((f"RU d{d} k1 C{c}", "conv", c, c, 1, 1, 1, B * T, {"skip": True, "stream_out": not last}))
    out.append(("tail k7", "tail", c, a["io"], 7, 1, 1, B * T, {}))
    return out


def model(layer):
    name, kind, Cin, Cout, K, s, d, rows, ex = layer
    if kind == "convT":
        flops = 2.0 * (rows / s) * Cin * Cout * K          # all taps of every input position
        taps_per_tile = (K + s - 1) // s                   # taps of one output phase
    else:
        flops = 2.0 * rows * Cin * Cout * K
        taps_per_tile = K
    if kind == "ru":
        flops += 2.0 * rows * Cin * Cout                   # the k=1 conv
    # algorithmic HBM bytes per output row
    if kind == "ru":
        hbm = 2 * Cin + 2 * Cin + (2 * Cout if ex.get("stream_out") else 0) + 2 * Cout
    elif kind == "tail":
        hbm = 2 * Cin + 4 * Cout
    else:
        rows_in = rows / s if kind == "convT" else rows
        hbm = 2 * Cin * rows_in / rows + 2 * Cout + (2 * Cout if ex.get("skip") else 0) + (2 * Cout if ex.get("stream_out") else 0)
    # one tile: 128 out-channels x 256 output rows of one phase (swap orientation), weights re-streamed per tile
    if kind == "tail" or Cout % 128 or Cin % 64:
        return dict(name=name, gflop=flops / 1e9, hbm_gb=hbm * rows / 1e9, ingest=None, mma_cyc=None)
    slab_rows = 256 + (K - 1) * d if kind != "convT" else 256 + taps_per_tile
    w_bytes = taps_per_tile * 128 * Cin * 2 + (128 * 128 * 2 * 2 if kind == "ru" else 0)
    a_bytes = slab_rows * Cin * 2 + (256 * Cout * 2 if (kind == "ru" or ex.get("skip")) else 0)
    tile_flops = 2.0 * 256 * 128 * Cin * taps_per_tile + (2.0 * 256 * 128 * 128 if kind == "ru" else 0)
    mma_cyc = tile_flops / MMA_FLOP_PER_CLK
    return dict(name=name, gflop=flops / 1e9, hbm_gb=hbm * rows / 1e9, ingest=(w_bytes + a_bytes) / mma_cyc,
                mma_cyc=mma_cyc, w_share=w_bytes / (w_bytes + a_bytes), pair=(w_bytes / 2 + a_bytes) / mma_cyc)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "sao"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    a = ARCH[which]
    rows = [model(l) for l in decoder_layers(a, B)]
    tg, th = sum(r["gflop"] for r in rows), sum(r["hbm_gb"] for r in rows)
    print(f"{which} decode, {B} clips x {a['frames']} latent frames: {tg / 1e3:.2f} TFLOP, {th:.1f} GB of algorithmic HBM traffic "
          f"({tg / th:.0f} FLOP/B); L2->SM ingest at full MMA rate vs ~{L2_TO_SM_CAP:.0f} B/clk/SM deliverable")
    print(f"{'launch':24s} {'GFLOP':>9s} {'HBM GB':>8s} {'FLOP/B':>7s} {'MMA cyc/tile':>13s} {'ingest B/clk':>13s} {'weights':>8s} {'max tensor %':>13s} {'CTA pair B/clk':>15s}")
    for r in rows:
        if r["ingest"] is None:
            print(f"{r['name']:24s} {r['gflop']:9.1f} {r['hbm_gb']:8.2f} {r['gflop'] / r['hbm_gb']:7.0f}")
            continue
        cap = min(1.0, L2_TO_SM_CAP / r["ingest"])
        print(f"{r['name']:24s} {r['gflop']:9.1f} {r['hbm_gb']:8.2f} {r['gflop'] / r['hbm_gb']:7.0f} {r['mma_cyc']:13.0f} "
              f"{r['ingest']:13.1f} {100 * r['w_share']:7.0f}% {100 * cap:12.0f}% {r['pair']:15.1f}")
    c128 = sum(r["hbm_gb"] for r in rows if "C128" in r["name"] or "->128" in r["name"] or r["name"].startswith("tail"))
    ru = [r for r in rows if r["name"].startswith("fused RU")]
    saved = sum(r["hbm_gb"] for r in ru) / 2
    print(f"128-channel stages hold {100 * c128 / th:.0f} % of the HBM bytes; a ResidualUnit that builds its operand from the stream "
          f"(512 instead of 1024 B per row) would save {saved:.1f} GB = {100 * saved / th:.0f} % of all HBM traffic")


if __name__ == "__main__":
    main()
