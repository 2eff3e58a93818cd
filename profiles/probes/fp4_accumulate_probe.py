"""Probe (library kernels, NOT a product path): does the Blackwell FP4 / FP8 tensor pipe accumulate small-integer
dot products of length K = 17792 exactly?  cuBLASLt through torch._scaled_mm, fp32 output, unit scales."""
import torch
torch.manual_seed(0)
M, N, K = 256, 512, 17792
a = torch.randint(0, 3, (M, K), device="cuda")            # {0,1,2}
a[: M // 2] = a[: M // 2].clamp(min=1)                    # dense half: sums beyond 2^14 stress the accumulator width
b = (torch.rand((N, K), device="cuda") < 0.9).long()      # {0,1}
want = (a.double() @ b.double().t())
print("max exact value", want.max().item())

def report(tag, got):
    d = (got.double() - want).abs()
    print("%s: max abs err %.1f, mismatching entries %d of %d" % (tag, d.max().item(), int((d != 0).sum()), d.numel()))

# ---- fp8 e4m3, tensorwise unit scales
try:
    a8, b8 = a.to(torch.float8_e4m3fn), b.to(torch.float8_e4m3fn)
    one = torch.ones((), device="cuda", dtype=torch.float32)
    for fast in (False, True):
        out = torch._scaled_mm(a8, b8.t(), scale_a=one, scale_b=one, out_dtype=torch.float32, use_fast_accum=fast)
        report("fp8 e4m3 fast_accum=%s" % fast, out)
except Exception as e:
    print("fp8 probe failed:", type(e).__name__, str(e)[:300])

# ---- mxfp4 (e2m1, e8m0 scale per 32): codes 0 -> 0b0000, 1.0 -> 0b0010, 2.0 -> 0b0100
try:
    code = torch.tensor([0, 2, 4], device="cuda", dtype=torch.uint8)
    def pack(x):
        c = code[x]
        return (c[:, 0::2] | (c[:, 1::2] << 4)).contiguous().view(torch.float4_e2m1fn_x2)
    a4, b4 = pack(a), pack(b)
    def unit_scales(rows):
        # nvfp4: one e4m3 scale per 16 elements; 1.0 everywhere (layout-independent because uniform)
        r = (rows + 127) // 128 * 128
        kb = (K // 16 + 3) // 4 * 4
        return torch.ones((r * kb,), dtype=torch.float32, device="cuda").to(torch.float8_e4m3fn)
    out = torch._scaled_mm(a4, b4.t(), scale_a=unit_scales(M), scale_b=unit_scales(N), out_dtype=torch.float32)
    report("nvfp4 e2m1 (block 16, unit e4m3 scales)", out)
except Exception as e:
    print("mxfp4 probe failed:", type(e).__name__, str(e)[:3000])
    try:
        out = torch._scaled_mm(a4, b4.t(), scale_a=unit_scales(M), scale_b=unit_scales(N), out_dtype=torch.bfloat16)
        print("mxfp4 bf16 output runs; sample", out[0, :4].float().tolist(), want[0, :4].tolist())
    except Exception as e2:
        print("mxfp4 bf16 also failed:", type(e2).__name__, str(e2)[:300])
