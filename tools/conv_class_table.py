"""Per-shape-class table of the convolution launches of one B=32 UNet forward from an ncu summary CSV
(tools/ncu_extract.py output, launches in forward order): FLOPs, time, TFLOP/s and tensor-pipe activity per class.
The launch order follows the planned forward (csrc/unet_engine.cu) of the reference topology:
   python tools/conv_class_table.py profiles/r02_v4_ncu_conv.csv > profiles/r02_v4_conv_classes.md"""
import collections
import csv
import sys

B = 32


def forward_order():
    """(label, H, Cout, K of the 3x3 part, K of a fused 1x1 skip) for the 56 conv_igemm2 launches, in launch order."""
    seq = []

    def res(h, cin, cout, skip_c=0, n=1):
        for _ in range(n):
            seq.append((f"{cin}->{cout} @{h}x{h}", h, cout, 9 * cin, 0))
            tag = f"{cout}->{cout} @{h}x{h} + " + (f"1x1 skip over {skip_c}" if skip_c else "identity skip")
            seq.append((tag, h, cout, 9 * cout, skip_c))

    res(96, 128, 128, n=3)            # input_blocks 1-3
    res(48, 128, 128)                 # 4 (down)
    res(48, 128, 256, skip_c=128)     # 5
    res(48, 256, 256, n=2)            # 6-7
    res(24, 256, 256)                 # 8 (down)
    res(24, 256, 256, n=3)            # 9-11
    res(24, 256, 256)                 # middle 0
    seq.append(("attention qkv 1x1 256->768 @24x24", 24, 768, 256, 0))
    seq.append(("attention proj_out 1x1 256->256 @24x24", 24, 256, 256, 0))
    res(24, 256, 256)                 # middle 2
    res(24, 512, 256, skip_c=512, n=4)  # output 0-3
    res(48, 256, 256)                 # output 3 up-sampling ResBlock
    res(48, 512, 256, skip_c=512, n=3)  # output 4-6
    res(48, 384, 256, skip_c=384)     # output 7
    res(96, 256, 256)                 # output 7 up-sampling ResBlock
    res(96, 384, 128, skip_c=384)     # output 8
    res(96, 256, 128, skip_c=256, n=3)  # output 9-11
    return seq


def main():
    rows = list(csv.DictReader(open(sys.argv[1])))
    seq = forward_order()
    assert len(seq) == 56
    cls = collections.OrderedDict()
    tot_t = tot_f = 0.0
    for r, (label, h, cout, k3, k1) in zip(rows, seq):
        fl = 2.0 * B * h * h * cout * (k3 + k1)
        t = float(r["time_us"])
        c = cls.setdefault(label, [0, 0.0, 0.0, 0.0, 0.0])
        c[0] += 1
        c[1] += t
        c[2] += fl
        c[3] += t * float(r["tensor_pipe_pct"])
        c[4] += float(r["dram_read_MB"]) + float(r["dram_write_MB"])
        tot_t += t
        tot_f += fl
    print(f"Convolution launches of one B={B} UNet forward by shape class ({sys.argv[1]}; ncu --set full: cold caches, "
          f"serialised launches, the box's clock under ncu - compare shares and tensor-pipe activity, not absolute times; "
          f"{len(rows)} of 56 launches captured)\n")
    print("| class | launches | K | GFLOP each | us each | TFLOP/s | tensor pipe active % | DRAM MB each | share of conv time % |")
    print("|---|---|---|---|---|---|---|---|---|")
    for label, (n, t, fl, tp, mb) in cls.items():
        k = next(k3 + k1 for (l, _, _, k3, k1) in seq if l == label)
        print(f"| {label} | {n} | {k} | {fl / n / 1e9:.1f} | {t / n:.1f} | {fl / t / 1e6:.0f} | {tp / t:.1f} | {mb / n:.0f} | "
              f"{100 * t / tot_t:.1f} |")
    print(f"\nTotal: {tot_t:.0f} us for {tot_f / 1e12:.3f} TFLOP = {tot_f / tot_t / 1e6:.0f} TFLOP/s under ncu.")


if __name__ == "__main__":
    main()
