"""Which fp32 arithmetic does F.batch_norm (eval, CUDA) use?  Compares candidate formulas bit for bit."""
import json, os, sys, torch
dev = "cuda:0"
torch.manual_seed(3)
res = {}
for shape in ((8, 64, 28, 28), (4, 512, 7, 7), (2, 3, 224, 224)):
    C = shape[1]
    x = torch.randn(shape, device=dev) * 3
    mean, var = torch.randn(C, device=dev), torch.rand(C, device=dev) + 0.3
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    eps = 1e-5
    ref = torch.nn.functional.batch_norm(x, mean, var, gamma, beta, False, 0.0, eps)
    v = lambda t: t.view(1, -1, 1, 1)
    d = lambda t: t.double()
    bits = lambda t: t.contiguous().view(torch.int32)
    ne = lambda a: int((bits(a) != bits(ref)).sum())
    rs = torch.rsqrt(var + eps)
    inv = 1.0 / torch.sqrt(var + eps)
    out = {}
    for nm, istd in (("rsqrt", rs), ("1/sqrt", inv)):
        t2 = (x - v(mean)) * v(istd)                       # two roundings
        out[f"{nm}: fma((x-m)*istd, g, b)"] = ne((d(t2) * d(v(gamma)) + d(v(beta))).float())
        out[f"{nm}: ((x-m)*istd)*g + b"] = ne(t2 * v(gamma) + v(beta))
        w = v(gamma * istd)
        out[f"{nm}: fma(x-m, g*istd, b)"] = ne((d(x - v(mean)) * d(w) + d(v(beta))).float())
        out[f"{nm}: fma(g*(x-m), istd, b)"] = ne((d(v(gamma) * (x - v(mean))) * d(v(istd)) + d(v(beta))).float())
        sc = gamma * istd
        sh = beta - mean * sc
        out[f"{nm}: fma(x, g*istd, b-m*g*istd)"] = ne((d(x) * d(v(sc)) + d(v(sh))).float())
        sh2 = (d(beta) - d(mean) * d(sc)).float()
        out[f"{nm}: fma(x, sc, fma(-m,sc,b))"] = ne((d(x) * d(v(sc)) + d(v(sh2))).float())
    out["n"] = x.numel()
    res[str(shape)] = out
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bn_formula.json", "w"), indent=1)
