"""Per-block timing of the decoder on the B200 (CUDA events, L2 flushed between iterations): the own tcgen05
convolution (mhada_conv3x3) next to the cuDNN route r1 used, and the pad kernels.  Development aid."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhada_style_transfer_b200 import network as N

flush = None
def t(fn, it=10):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2): fn()
    ms = []
    for _ in range(it):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return min(ms)

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    hw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dec = N.Decoder().cuda().eval()
    x = torch.randn(B, hw, hw, 512, device="cuda").bfloat16()
    rows, tot_own, tot_lib = [], 0.0, 0.0
    with torch.no_grad():
        whole = t(lambda: dec(x.permute(0, 3, 1, 2)))
        up = False
        for i, blk in enumerate(dec._blocks()):
            cin, cout = blk.conv.conv.in_channels, blk.conv.conv.out_channels
            w, b = blk.conv._weights(x.dtype)
            if cin == 64 and cout <= 8:
                ms = t(lambda: N._conv3x3_small_relu(x, blk.conv.conv.weight, blk.conv.conv.bias))
                rows.append({"block": i, "shape": f"{cin}->{cout}@{x.shape[1]}x{x.shape[2]}", "own_small_us": round(ms * 1e3, 1)})
                tot_own += ms; tot_lib += ms
                break
            xp = N._pad_reflect(x, up)
            ms_pad = t(lambda: N._pad_reflect(x, up))
            ms_lib = t(lambda: N._conv3x3_relu(xp.permute(0, 3, 1, 2), w, b))
            wp, bp = blk.conv._packed_tc()
            ms_own = t(lambda: N._conv3x3_tc_relu(xp, wp, bp, False))
            ms_own_p = t(lambda: N._conv3x3_tc_relu(xp, wp, bp, True))
            y = N._conv3x3_tc_relu(xp, wp, bp, False)
            H, W = xp.shape[1] - 2, xp.shape[2] - 2
            fl = 2 * 9 * cin * cout * H * W * B
            rows.append({"block": i, "shape": f"{cin}->{cout}@{H}x{W}", "pad_us": round(ms_pad * 1e3, 1), "up": up,
                         "cudnn_us": round(ms_lib * 1e3, 1), "cudnn_tflops": round(fl / ms_lib / 1e9),
                         "own_us": round(ms_own * 1e3, 1), "own_tflops": round(fl / ms_own / 1e9),
                         "own_padded_out_us": round(ms_own_p * 1e3, 1)})
            tot_lib += ms_pad + ms_lib
            tot_own += ms_own + (ms_pad if (up or i == 0) else 0.0)
            x = y
            up = bool(blk.scale_factor)
    print(json.dumps({"B": B, "hw": hw, "decoder_forward_ms": round(whole, 4), "sum_own_ms": round(tot_own, 4),
                      "sum_r1_route_ms": round(tot_lib, 4), "blocks": rows}))

if __name__ == "__main__":
    main()
