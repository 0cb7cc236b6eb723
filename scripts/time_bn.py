"""Device time of the BatchNorm-block / LayerNorm kernels on the buffers of a B=256, F=270 engine (left stream, encoder 1), each
call replayed 10x from a CUDA graph (no host gaps, no overlap with other kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, out_dim, T_LEN
from multi_modal_csi_b200 import THAT, FusedAdam
from multi_modal_csi_b200 import layout as LY
from multi_modal_csi_b200.engine import P_DROP, BN_EPS, BN_MOMENTUM, LN_EPS
from multi_modal_csi_b200.layout import HALO, site

F = int(sys.argv[1]) if len(sys.argv) > 1 else 270
B = 256
dev = torch.device("cuda", 0)
torch.manual_seed(39)
model = THAT((T_LEN, F), (out_dim(F),), act_dtype="bf16", max_batch=B).to(dev)
model.train()
opt = FusedAdam(model.parameters(), lr=5e-4, weight_decay=2e-4)
x, y = synth_batch(B, F, out_dim(F), 1234)
x, y = x.to(dev), y.to(dev)
for _ in range(2):
    model.fused_train_step(x, y, opt, pos_weight=4.0, augment=True, use_graph=False)
torch.cuda.synchronize()
eng = model._engine
ops = eng.ops


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for si in (0, 1):
    sg = eng.g.streams[si]
    st = eng.s[sg.name]
    e = 1 if sg.n_enc > 1 else 0
    a, p = st["enc"][e], sg.prefix(e)
    d, Dp, L = sg.d, sg.Dp, sg.L
    n = B * L * d
    gam, bet = eng._bn3(sg, e, "1.weight"), eng._bn3(sg, e, "1.bias")
    sb, so = site(si, e, LY.SITE_BRANCH), site(si, e, LY.SITE_SUM)
    dout, dz = st["dout"][0], st["dz"][e]
    res = {}
    res["bn_stats"] = (timeit(lambda: ops.bn_stats(a["z"].t, B, L, HALO, 3 * Dp, a["bn_sums"])), n * 6)
    res["bn_act_fwd"] = (timeit(lambda: ops.bn_act_fwd(a["z"].t, a["bn_mean"], a["bn_invstd"], gam, bet, a["t"].t, a["out"].t, B, L, d, HALO, 3,
                                                       P_DROP, sb, P_DROP, so, eng.rng, a["dmask"])), n * 14)
    res["bn_act_bwd_reduce"] = (timeit(lambda: ops.bn_act_bwd_reduce(dout.t, a["z"].t, a["bn_mean"], a["bn_invstd"], gam, bet, B, L, d, HALO, 3,
                                                                     P_DROP, sb, P_DROP, so, eng.rng, a["red"], a["dmask"])), n * 10)
    res["bn_act_bwd_dz"] = (timeit(lambda: ops.bn_act_bwd_dz(dout.t, a["z"].t, a["bn_mean"], a["bn_invstd"], gam, bet, a["red"], B, L, d, HALO, 3,
                                                             P_DROP, sb, P_DROP, so, eng.rng, dz.t, eng._bn3(sg, e, "1.weight", grad=True),
                                                             eng._bn3(sg, e, "1.bias", grad=True), a["dmask"])), n * 16)
    res["ln_fwd"] = (timeit(lambda: ops.layernorm_fwd(a["t"].t, eng.P(p + "layer_norm_1.weight"), eng.P(p + "layer_norm_1.bias"), a["s"].t,
                                                      a["mean1"], a["rstd1"], B, L, d, HALO, LN_EPS)), n * 6)
    res["ln_bwd(+masked copy)"] = (timeit(lambda: ops.layernorm_bwd(st["ds"].t, a["t"].t, eng.P(p + "layer_norm_1.weight"), a["mean1"], a["rstd1"], dout.t,
                                                                    st["dt"].t, st["dtm"][e].t, P_DROP, site(si, e, LY.SITE_ATTN), eng.rng,
                                                                    eng.G(p + "layer_norm_1.weight"), eng.G(p + "layer_norm_1.bias"), B, L, d, HALO)), n * 16)
    res["ln_bwd"] = (timeit(lambda: ops.layernorm_bwd(st["dt0"].t, st["x0"].t, eng.P(p + "layer_norm_0.weight"), a["mean0"], a["rstd0"], st["dt"].t,
                                                      dout.t, None, 0.0, 0, eng.rng, eng.G(p + "layer_norm_0.weight"),
                                                      eng.G(p + "layer_norm_0.bias"), B, L, d, HALO)), n * 14)
    for k, (us, nb) in res.items():
        print(f"{sg.name:5s} {k:22s} {us:7.1f} us  {nb / us / 1e3:7.0f} GB/s", flush=True)
