import torch, time
x = torch.randn(32, 1936, 128, device="cuda")
l1 = torch.nn.Linear(128, 128, bias=False).cuda(); mlp1 = torch.nn.Linear(256, 1024, bias=False).cuda(); mlp2 = torch.nn.Linear(1024, 128, bias=False).cuda()
ln = torch.nn.LayerNorm(128).cuda(); gelu = torch.nn.GELU()
def t(fn, it=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it * 1e3
with torch.no_grad():
    print("allow_tf32:", torch.backends.cuda.matmul.allow_tf32)
    print("Linear 128->128 on [32,1936,128]: %.1f us" % t(lambda: l1(x)))
    xc = torch.cat([x, x], -1)
    print("Linear 256->1024: %.1f us" % t(lambda: mlp1(xc)))
    h = mlp1(xc)
    print("GELU on [32,1936,1024]: %.1f us" % t(lambda: gelu(h)))
    print("Linear 1024->128: %.1f us" % t(lambda: mlp2(h)))
    print("LayerNorm(128): %.1f us" % t(lambda: ln(x)))
    print("cat: %.1f us" % t(lambda: torch.cat([x, x], -1)))
