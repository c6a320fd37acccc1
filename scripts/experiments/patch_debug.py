import sys, torch
sys.path.insert(0, ".")
from a3vlm_b200 import ops
bf16 = torch.bfloat16
B, S, P, D = 2, 224, 14, 1024
g = torch.Generator(device="cuda").manual_seed(1)
img = torch.randn(B, 3, S, S, device="cuda", generator=g)
w = (torch.randn(D, 3, P, P, device="cuda", generator=g) * 0.05).to(bf16)
grid = S // P
img16 = torch.zeros(B, 3, S, grid, 16, device="cuda", dtype=bf16)
ops.vit_pad_cast(img, img16, B, S, P)
print("pad ok", torch.equal(img16[..., :P].reshape(B, 3, S, S), img.to(bf16)))
w768 = torch.zeros(D, 3, 16, 16, device="cuda", dtype=bf16); w768[:, :, :P, :P] = w
out = torch.full((B * grid * grid, D), float("nan"), device="cuda", dtype=bf16)
ops.vit_patch_embed(img16, w768.view(D, 768), out, B, S, P)
ref = torch.nn.functional.conv2d(img.to(bf16).float(), w.float(), stride=P).permute(0, 2, 3, 1).reshape(B * grid * grid, D)
o = out.float()
print("nan rows", int(torch.isnan(o).any(1).sum()), "of", o.shape[0])
err = (o - ref).abs()
print("max err", err.max().item(), "ref absmax", ref.abs().max().item())
# is it a row permutation?  match each output row to the nearest ref row
o2 = torch.nan_to_num(o)
d = torch.cdist(o2[:64], ref)
print("row 0..15 nearest ref rows:", d.argmin(1)[:16].tolist(), "dist", [round(x, 3) for x in d.min(1).values[:8].tolist()])
# per-channel / per-ky contribution check: use weights that pick one (c, ky, kx)
for (c, ky, kx) in ((0, 0, 0), (0, 5, 3), (1, 13, 13), (2, 12, 0), (0, 0, 13)):
    w1 = torch.zeros(D, 3, 16, 16, device="cuda", dtype=bf16); w1[0, c, ky, kx] = 1
    ops.vit_patch_embed(img16, w1.view(D, 768), out, B, S, P)
    want = img.to(bf16)[:, c].reshape(B, grid, P, grid, P)[:, :, ky, :, kx].reshape(-1).float()
    got = out[:, 0].float()
    print((c, ky, kx), "match", torch.equal(got, want), "first got", got[:4].tolist(), "want", want[:4].tolist())
