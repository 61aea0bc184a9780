"""Isolating probes for the IMMA decode-attention kernel (development aid; run on a B200):
   A1  K = 0, V depends on the channel only      -> output = V row           (channel mapping of the V phase / flush)
   A2  K = 0, V random                            -> output = mean_t V        (token <-> k-slot mapping, transposes, weight planes)
   B   V one-hot in one token, K random           -> output = softmax weight  (K phase: ldmatrix rows, swizzle, score assembly)
   C   random everything                          -> full comparison
"""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-llm-inference_b200"))
import kvq_b200  # noqa: E402


def pack4(codes):  # int codes in [-8, 7], [..., D] -> uint8 [..., D/2]
    u = (codes + 8).to(torch.uint8)
    return (u[..., 0::2] << 4) | u[..., 1::2]


def ref(q, kcodes, ks, vcodes, vs, G):
    B, H, T, D = kcodes.shape
    kd = kcodes.double() * ks.double()[None, None, :, None]
    vd = vcodes.double() * vs.double()[None, None, :, None]
    qf = q[:, :, 0].double().view(B, H, G, D)
    p = torch.softmax(torch.einsum("bhgd,bhtd->bhgt", qf, kd) * D ** -0.5, dim=-1)
    return torch.einsum("bhgt,bhtd->bhgd", p, vd).reshape(B, H * G, 1, D).float()


def run(kb, vb, G, D, T, probe, seed=0, B=1, H=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    klo, khi = (-127, 128) if kb == 8 else (-7, 8)
    vlo, vhi = (-127, 128) if vb == 8 else (-7, 8)
    kcodes = torch.randint(klo, khi, (B, H, T, D), device="cuda", generator=g)
    vcodes = torch.randint(vlo, vhi, (B, H, T, D), device="cuda", generator=g)
    ks = torch.rand(T, device="cuda", generator=g) * 0.01 + 0.01
    vs = torch.rand(T, device="cuda", generator=g) * 0.01 + 0.01
    q = torch.randn(B, H * G, 1, D, device="cuda", generator=g)
    if probe in ("A1", "A2"):
        kcodes.zero_()
    if probe == "A1":
        vcodes[:] = vcodes[:, :, :1]
        vs[:] = 0.0125
    if probe == "R":  # scores ramp upwards with t: the softmax reference moves in (almost) every tile
        kcodes[:] = (q[:, :1, 0, :].sign() * 100).long()[:, :, None, :].expand(B, H, T, D) if G == 1 else kcodes
        if G > 1:
            kcodes[:] = (q[:, ::G, 0, :].sign() * 100).long()[:, :, None, :]
        ks[:] = torch.linspace(0.0005, 0.02, T, device="cuda")
    if probe == "O":  # outliers
        kcodes = torch.where(torch.rand(kcodes.shape, device="cuda", generator=g) < 0.01, kcodes.sign() * khi - kcodes.sign(), kcodes // 8)
    if probe == "B":
        t0 = T // 3
        vcodes.zero_()
        vcodes[:, :, t0] = 1
    kc = kcodes.to(torch.int8) if kb == 8 else pack4(kcodes)
    vc = vcodes.to(torch.int8) if vb == 8 else pack4(vcodes)
    want = ref(q, kcodes, ks, vcodes, vs, G)
    out = {}
    for variant in ("cuda_core", "imma"):
        got = kvq_b200.decode_attention_quantized(q, kc, ks, kb, vc, vs, vb, out_dtype=torch.float32, variant=variant)
        out[variant] = float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))
    return out, want


if __name__ == "__main__":
    torch.manual_seed(0)
    cases = [(8, 8, 1, 128), (8, 4, 1, 128), (8, 4, 8, 128), (4, 4, 4, 128), (8, 8, 2, 64), (4, 4, 1, 64), (4, 8, 8, 64), (8, 4, 4, 64)]
    if len(sys.argv) > 1:
        cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    Ts = (32, 128, 1000) if len(sys.argv) <= 1 else (640, 5000, 40000)
    probes = ("A1", "A2", "B", "C") if len(sys.argv) <= 1 else ("C", "R", "O")
    for (kb, vb, G, D), T, probe in itertools.product(cases, Ts, probes):
        try:
            err, _ = run(kb, vb, G, D, T, probe, H=4 if len(sys.argv) > 1 else 1)
            flag = "ok " if err["imma"] < 1e-4 else "BAD"
            print(f"{flag} K{kb} V{vb} G{G} D{D} T{T:5d} {probe}: imma {err['imma']:.3e}  cuda_core {err['cuda_core']:.3e}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"EXC K{kb} V{vb} G{G} D{D} T{T} {probe}: {type(e).__name__}: {str(e)[:200]}", flush=True)
            if "CUDA" in str(e):
                break
