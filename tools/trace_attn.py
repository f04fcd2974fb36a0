"""Timeline of one CTA of the tcgen05 attention kernel (clock64 stamps, mhada_debug_attn_trace)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhada_style_transfer_b200 import _lib

def main():
    L = _lib.lib(); dev = "cuda:0"
    B, H, d, C = 8, 8, 64, 512
    Nc = Ns = 4096
    g = torch.Generator(device=dev).manual_seed(0)
    q = (torch.randn(B, Nc, C, device=dev, generator=g) * 0.6).bfloat16()
    k = (torch.randn(B, Ns, C, device=dev, generator=g) * 0.6).bfloat16()
    v = (torch.randn(B, Ns, 2 * C, device=dev, generator=g) * 40).bfloat16()
    x = (torch.randn(B, Nc, C, device=dev, generator=g) * 30).bfloat16()
    st = torch.zeros(3, B, C, device=dev); st[1] = 1.0
    out = torch.empty_like(x)
    trace = torch.zeros(4 * 64 * 8 + 3 * 4096, dtype=torch.int64, device=dev)
    a = _lib.AttnArgs(); a.dtype = _lib.BF16
    a.B, a.H, a.Nc, a.Ns, a.dqk, a.dv = B, H, Nc, Ns, d, d
    a.q, a.k, a.v, a.x, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), x.data_ptr(), out.data_ptr()
    a.ldq, a.ldk, a.ldv, a.ldx, a.ldo = C, C, 2 * C, C, C
    a.x_mean, a.x_rstd, a.mu_v = st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        _lib.check("trace", L.mhada_debug_attn_trace(ctypes.byref(a), ctypes.c_void_p(trace.data_ptr()), stream))
    torch.cuda.synchronize()
    full = trace.cpu().numpy()
    t = full[:4 * 64 * 8].reshape(4, 64, 8)
    cta = full[4 * 64 * 8:].reshape(-1, 3)
    cta = cta[cta[:, 2] > 0]
    if len(cta):
        dur = (cta[:, 2] - cta[:, 1]) / 1e3
        start = (cta[:, 1] - cta[:, 1].min()) / 1e3
        end = (cta[:, 2] - cta[:, 1].min()) / 1e3
        print(f'CTAs {len(cta)}: duration us min {dur.min():.1f} median {np.median(dur):.1f} max {dur.max():.1f}; start spread {start.max():.1f} us; kernel span {end.max():.1f} us')
        order = np.argsort(dur)
        print('slowest SMs', [(int(cta[i,0]), round(float(dur[i]),1)) for i in order[-6:]], 'fastest', [(int(cta[i,0]), round(float(dur[i]),1)) for i in order[:6]])
    t0 = t[0, 0, 0]
    T = Ns // 64
    print("softmax WG0/WG1: [S ready, S in regs, max done, P st issued, P arrived]; MMA: [P0 seen, PV0+S0 issued, P1 seen, PV1+S1 issued]  (cycles rel. to first S0 ready)")
    for j in list(range(0, 8)) + list(range(T - 3, T)):
        r = lambda role, n: " ".join(f"{int(x - t0):7d}" for x in t[role, j, :n])
        print(f"j={j:2d} WG0 {r(0,5)} | WG1 {r(1,5)} | MMA {r(2,6)} | TMA {r(3,3)}")
    entry, exit_ = t[3, 0, 4], t[3, 0, 5]
    print("CTA life", int(exit_ - entry), "cycles: entry->first S ready", int(t0 - entry), "| last P arrive -> O ready",
          int(t[0, 63, 5] - t[0, T - 1, 4]), "| epilogue", int(t[0, 63, 6] - t[0, 63, 5]), "| epilogue end -> exit", int(exit_ - t[0, 63, 6]),
          "| mainloop", int(t[0, T - 1, 4] - t0))
    d0 = np.diff(t[0, :T, 0]); d1 = np.diff(t[1, :T, 0])
    print("period WG0 mean", d0[2:].mean(), "WG1", d1[2:].mean())
    for name, role in (("WG0", 0), ("WG1", 1)):
        seg = t[role, 2:T - 1]
        print(name, "ld", (seg[:, 1] - seg[:, 0]).mean(), "max", (seg[:, 2] - seg[:, 1]).mean(), "exp+st", (seg[:, 3] - seg[:, 2]).mean(),
              "wait_st+arrive", (seg[:, 4] - seg[:, 3]).mean(), "S-wait", (t[role, 3:T, 0] - t[role, 2:T - 1, 4]).mean())
    m = t[2, 2:T - 1]
    print("MMA: P0seen->issued", (m[:, 1] - m[:, 0]).mean(), "issued0->P1seen", (m[:, 2] - m[:, 1]).mean(), "P1seen->issued", (m[:, 3] - m[:, 2]).mean(),
          "issued1->next P0 seen", (t[2, 3:T, 0] - m[:, 3]).mean())
    print("P0 arrive -> MMA sees P0", (t[2, 2:T - 1, 0] - t[0, 2:T - 1, 4]).mean(), " P1 arrive -> MMA sees P1", (t[2, 2:T - 1, 2] - t[1, 2:T - 1, 4]).mean())

if __name__ == "__main__":
    main()
