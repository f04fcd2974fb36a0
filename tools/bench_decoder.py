"""Per-block timing of the decoder on the B200 (CUDA events, L2 flushed between iterations). Development aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhada_style_transfer_b200 import network as N

def t(fn, it=10):
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
    dec = N.Decoder().cuda().eval()
    x = torch.randn(B, 64, 64, 512, device="cuda").bfloat16()
    total = 0.0
    with torch.no_grad():
        up = False
        for i, blk in enumerate(dec._blocks()):
            cin, cout = blk.conv.conv.in_channels, blk.conv.conv.out_channels
            w, b = blk.conv._weights(x.dtype)
            if cin == 64 and cout <= 8:
                ms = t(lambda: N._conv3x3_small_relu(x, blk.conv.conv.weight, blk.conv.conv.bias))
                xp = N._pad_reflect(x, up)
                ms_pad = t(lambda: N._pad_reflect(x, up)); ms_conv = t(lambda: N._conv3x3_relu(xp.permute(0, 3, 1, 2), w, b))
                print(f"block {i}: {cin}->{cout} @ {x.shape[1]}x{x.shape[2]}: own fused {ms*1e3:.1f} us | library route pad {ms_pad*1e3:.1f} + conv {ms_conv*1e3:.1f} us")
                total += ms
                break
            xp = N._pad_reflect(x, up)
            ms_pad = t(lambda: N._pad_reflect(x, up))
            ms_conv = t(lambda: N._conv3x3_relu(xp.permute(0, 3, 1, 2), w, b))
            y = N._conv3x3_relu(xp.permute(0, 3, 1, 2), w, b)
            H, W = xp.shape[1] - 2, xp.shape[2] - 2
            fl = 2 * 9 * cin * cout * H * W * B
            print(f"block {i}: {cin}->{cout} @ {H}x{W}: pad{'+up' if up else ''} {ms_pad*1e3:.1f} us ({(x.numel()+xp.numel())*2/ms_pad/1e6:.0f} GB/s) | conv {ms_conv*1e3:.1f} us ({fl/ms_conv/1e9:.0f} TFLOP/s)")
            total += ms_pad + ms_conv
            x = N._token_major(y, y.dtype)
            up = bool(blk.scale_factor)
    print(f"decoder total {total:.3f} ms for batch {B}")

if __name__ == "__main__":
    main()
