import sys, math, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from dgdm_histopath_lab_b200 import ops, synth, train as T
from dgdm_histopath_lab_b200.graph import get_context
from oracle import dgdm_oracle as O
from util import SMALL_CFG, make_model, rel_err
DEV = "cuda"
m = make_model(SMALL_CFG).eval()
sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in m.state_dict().items()}
mg = make_model(SMALL_CFG).to(DEV).eval()
d = synth.tile_graph(300, 32, seed=3)
g = torch.Generator().manual_seed(0)
go = torch.randn(300, 32, generator=g)

def report(tag, prefix):
    bad = []
    for n, p in mg.named_parameters():
        if not n.startswith(prefix): continue
        gr = sd[n].grad
        if gr is None or float(gr.abs().max()) == 0: continue
        e = rel_err(p.grad, gr)[0] if p.grad is not None else float('nan')
        bad.append((n[len(prefix):], round(e, 6)))
    print(tag, bad)
    for v in sd.values():
        if v.is_floating_point(): v.grad = None
    mg.zero_grad()

# 1. dyn layer at level 0
x = d.x.clone().requires_grad_(True)
ref = O.dynamic_graph_layer(sd, "hierarchical_processor.down_convs.0", x, d.edge_index, None)
ref.backward(go)
ctx = get_context(d.to(DEV))
lvl = T._Level(ctx.csr, None)
xg = d.x.to(DEV).requires_grad_(True)
out = T._dyn_layer(mg.hierarchical_processor.down_convs[0], xg, lvl, False)
print("dyn fwd", rel_err(out, ref)[0])
out.backward(go.to(DEV))
print("dyn dx", rel_err(xg.grad, x.grad)[0])
report("dyn layer", "hierarchical_processor.down_convs.0.")

# 2. conv only
x = d.x.clone().requires_grad_(True)
ref = O.graph_convolution(sd, "hierarchical_processor.down_convs.0.graph_conv1", x, d.edge_index, None)
ref.backward(go)
xg = d.x.to(DEV).requires_grad_(True)
out = T._conv(mg.hierarchical_processor.down_convs[0].graph_conv1, xg, lvl)
out.backward(go.to(DEV))
print("conv fwd", rel_err(out, ref)[0], "dx", rel_err(xg.grad, x.grad)[0])
report("conv", "hierarchical_processor.down_convs.0.graph_conv1.")

# 3. unet
x = d.x.clone().requires_grad_(True)
ref = O.graph_unet(sd, "hierarchical_processor", x, d.edge_index, None)
ref.backward(go)
xg = d.x.to(DEV).requires_grad_(True)
out = T._unet(mg.hierarchical_processor, xg, ctx, None, False)
print("unet fwd", rel_err(out, ref)[0])
out.backward(go.to(DEV))
print("unet dx", rel_err(xg.grad, x.grad)[0])
report("unet", "hierarchical_processor.")

# 4. spatial attention
x = d.x.clone().requires_grad_(True)
ref, _ = O.spatial_attention(sd, "spatial_attention", x, d.pos, 2)
ref.backward(go)
xg = d.x.to(DEV).requires_grad_(True)
out = T._spatial_attention(mg.spatial_attention, xg, d.pos.to(DEV), ctx, False)
print("attn fwd", rel_err(out, ref)[0])
out.backward(go.to(DEV))
print("attn dx", rel_err(xg.grad, x.grad)[0])
report("attn", "spatial_attention.")

# ---- batch of 2 x 150
print("==== batch 2x150")
b = synth.tile_batch(2, 150, 48, seed=11)
ctxb = get_context(b.to(DEV))
go = torch.randn(300, 32, generator=g)
x = torch.randn(300, 32, generator=g)
xr = x.clone().requires_grad_(True)
parts = []
for i in range(2):
    mk = b.batch == i
    o, _ = O.spatial_attention(sd, "spatial_attention", xr[mk], b.pos[mk], 2)
    parts.append(o)
ref = torch.cat(parts)
ref.backward(go)
xg = x.to(DEV).requires_grad_(True)
out = T._spatial_attention(mg.spatial_attention, xg, b.pos.to(DEV), ctxb, False)
print("attn(batch) fwd", rel_err(out, ref)[0])
out.backward(go.to(DEV))
print("attn(batch) dx", rel_err(xg.grad, xr.grad)[0])
report("attn(batch)", "spatial_attention.")

# encoder chain
xr = b.x.clone().requires_grad_(True)
h = O.feature_encoder(sd, "feature_encoder", xr)
ref = O.graph_encoder(sd, "graph_encoder", h, b.edge_index, None, 4)
ref.backward(go)
xg = b.x.to(DEV).requires_grad_(True)
lvl0 = T._Level(ctxb.csr, None)
out = T._graph_encoder(mg.graph_encoder, T._feature_encoder(mg.feature_encoder, xg, False), lvl0, False)
print("enc fwd", rel_err(out, ref)[0])
out.backward(go.to(DEV))
print("enc dx", rel_err(xg.grad, xr.grad)[0])
report("enc", "graph_encoder.graph_layers.0")

print("==== full model")
from util import golden, oracle_cfg
gd = golden("pretrain_step.npz")
rng = dict(mask_indices=torch.tensor(gd["mask_indices"]), mask_token=torch.tensor(gd["mask_token"]),
           timesteps=torch.tensor(gd["timesteps"]), noise=list(torch.tensor(gd["noise"]).split(150)),
           target=list(torch.tensor(gd["target"]).split(150)))
xm = O.apply_entity_masking(b.x, rng["mask_indices"], rng["mask_token"])
ref = O.dgdm_forward(sd, oracle_cfg(SMALL_CFG), xm, b.edge_index, None, b.batch, b.pos, mode="pretrain", rng=rng, return_stages=True)
from dgdm_histopath_lab_b200 import Data
bd = Data(x=xm.to(DEV), edge_index=b.edge_index.to(DEV), pos=b.pos.to(DEV)); bd.batch = b.batch.to(DEV)
rng_g = dict(rng, noise=torch.tensor(gd["noise"]), target=torch.tensor(gd["target"]))
out = T.train_forward(mg, bd, "pretrain", True, rng_g)
print("node_embeddings fwd", rel_err(out["node_embeddings"], ref["node_embeddings"])[0], "loss", float(out["diffusion_loss"]), float(ref["diffusion_loss"]))
# gradient of a plain functional of the node embeddings (no diffusion): isolates encoder+attention+unet
go = torch.randn(300, 32, generator=g)
(ref["node_embeddings"] * go).sum().backward(retain_graph=True)
(out["node_embeddings"] * go.to(DEV)).sum().backward(retain_graph=True)
bad = []
for n, p in mg.named_parameters():
    gr = sd[n].grad
    if gr is None or float(gr.abs().max()) < 1e-6 or p.grad is None: continue
    e = rel_err(p.grad, gr)[0]
    if e > 5e-4: bad.append((n, round(e, 4)))
print("h-functional: bad", len(bad), bad[:6])
for v in sd.values():
    if v.is_floating_point(): v.grad = None
mg.zero_grad()
ref["diffusion_loss"].backward()
out["diffusion_loss"].backward()
bad = []
for n, p in mg.named_parameters():
    gr = sd[n].grad
    if gr is None or float(gr.abs().max()) < 1e-9 or p.grad is None: continue
    e = rel_err(p.grad, gr)[0]
    if e > 5e-4: bad.append((n, round(e, 4), float(gr.abs().max())))
print("loss: bad", len(bad), bad[:6])
