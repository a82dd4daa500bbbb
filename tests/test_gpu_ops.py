"""Operator-level parity (-m gpu): every C-ABI operator against a plain fp32 PyTorch/oracle statement of the same op."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import O, assert_fp32, metrics

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(cuda):
    from hsenet_b200 import _lib
    return _lib.load()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _check(rc):
    from hsenet_b200 import _lib
    _lib.check(rc, "test")


# ---------------------------------------------------------------- integer maps: bit exact
def test_patch_gather_map_bit_exact(lib, cuda):
    out = torch.empty(2048, 1024, dtype=torch.int32, device=cuda)
    _check(lib.hsenet_patch_gather_map(out.data_ptr(), _st()))
    assert np.array_equal(out.cpu().numpy(), O.patch_gather_map())


def test_packer_window_map_bit_exact(lib, cuda):
    out = torch.empty(128, 16, dtype=torch.int32, device=cuda)
    _check(lib.hsenet_packer_window_map(out.data_ptr(), _st()))
    assert np.array_equal(out.cpu().numpy(), O.packer_window_map())


@pytest.mark.parametrize("B", [1, 3])
def test_im2col_bit_exact(lib, cuda, B):
    from hsenet_b200 import _lib
    x = torch.rand(B, 1, 32, 256, 256, generator=torch.Generator().manual_seed(B))
    ref = O.patchify(x)
    xd = x.to(cuda)
    out = torch.empty(B * 2048, 1024, dtype=torch.float32, device=cuda)
    _check(lib.hsenet_patch_im2col(xd.data_ptr(), B, out.data_ptr(), _lib.DTYPE_F32, _st()))
    assert torch.equal(out.cpu().reshape(B, 2048, 1024), ref)            # pure data movement: bit exact
    outb = torch.empty(B * 2048, 1024, dtype=torch.bfloat16, device=cuda)
    _check(lib.hsenet_patch_im2col(xd.data_ptr(), B, outb.data_ptr(), _lib.DTYPE_BF16, _st()))
    assert torch.equal(outb.cpu().reshape(B, 2048, 1024), ref.to(torch.bfloat16))   # RNE cast: bit exact


# ---------------------------------------------------------------- GEMM
def _gemm(lib, A, W, bias=None, resid=None, gelu=False, prec=0, want_f32=True, want_act=True):
    from hsenet_b200 import _lib
    M, K = A.shape
    N = W.shape[0]
    act = torch.bfloat16 if prec == 0 else torch.float32
    of = torch.full((M, N), float("nan"), dtype=torch.float32, device=A.device) if want_f32 else None
    oa = torch.full((M, N), float("nan"), dtype=act, device=A.device) if want_act else None
    rc = lib.hsenet_linear(A.data_ptr(), K, W.data_ptr(), K, M, N, K,
                           None if bias is None else bias.data_ptr(),
                           None if resid is None else resid.data_ptr(), N, int(gelu),
                           None if of is None else of.data_ptr(), N,
                           None if oa is None else oa.data_ptr(), N, prec, _st())
    _lib.check(rc, "linear")
    torch.cuda.synchronize()
    return of, oa


@pytest.fixture(params=["2cta", "1cta"])
def gemm_variant(request, monkeypatch):
    """The dispatcher reads HSENET_GEMM_1CTA per call: exercise the CTA-pair kernel and the 1-CTA kernel."""
    monkeypatch.setenv("HSENET_GEMM_1CTA", "1" if request.param == "1cta" else "0")
    return request.param


GEMM_SHAPES = [(128, 256, 64), (300, 256, 128), (2049, 768, 768), (4098, 2304, 768), (1000, 768, 3072),
               (64, 1536, 768), (2, 768, 768), (4096, 768, 1024), (16392, 768, 768), (257, 512, 64),
               (20000, 256, 3072)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bf16_plain(lib, cuda, gemm_variant, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    ref = A.float() @ W.float().T
    of, oa = _gemm(lib, A.to(cuda), W.to(cuda))
    m = metrics(of, ref)
    assert m["max_rel"] < 2e-5, m                  # bf16 products are exact in fp32; only summation order differs
    assert torch.equal(oa.cpu(), of.cpu().to(torch.bfloat16))


@pytest.mark.parametrize("M,N,K", [(2049, 768, 768), (513, 3072, 768)])
def test_gemm_bf16_epilogues(lib, cuda, gemm_variant, M, N, K):
    g = torch.Generator().manual_seed(11)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g)
    base = A.float() @ W.float().T + bias
    of, _ = _gemm(lib, A.to(cuda), W.to(cuda), bias=bias.to(cuda))
    assert metrics(of, base)["max_rel"] < 2e-5
    of, oa = _gemm(lib, A.to(cuda), W.to(cuda), bias=bias.to(cuda), gelu=True)
    assert metrics(of, F.gelu(base))["max_rel"] < 2e-5
    assert metrics(oa, F.gelu(base))["max_rel"] < 5e-3
    rd = resid.to(cuda)
    of, _ = _gemm(lib, A.to(cuda), W.to(cuda), bias=bias.to(cuda), resid=rd)
    assert metrics(of, base + resid)["max_rel"] < 2e-5


def test_gemm_bf16_inplace_residual(lib, cuda, gemm_variant):
    from hsenet_b200 import _lib
    M, N, K = 1500, 768, 768
    g = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).to(cuda)
    x = torch.randn(M, N, generator=g).to(cuda)
    ref = x.cpu() + A.float().cpu() @ W.float().cpu().T
    _lib.check(lib.hsenet_linear(A.data_ptr(), K, W.data_ptr(), K, M, N, K, None, x.data_ptr(), N, 0, x.data_ptr(), N,
                                 None, 0, 0, _st()), "linear")
    assert metrics(x, ref)["max_rel"] < 2e-5


def test_gemm_rejects_bad_shapes(lib, cuda):
    A = torch.zeros(16, 64, dtype=torch.bfloat16, device=cuda)
    W = torch.zeros(100, 64, dtype=torch.bfloat16, device=cuda)
    o = torch.zeros(16, 100, device=cuda)
    rc = lib.hsenet_linear(A.data_ptr(), 64, W.data_ptr(), 64, 16, 100, 64, None, None, 0, 0, o.data_ptr(), 100, None,
                           0, 0, _st())
    assert rc == -1


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (2049, 768, 768), (130, 3072, 768)])
def test_gemm_f32(lib, cuda, M, N, K):
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    bias = torch.randn(N, generator=g)
    ref = F.gelu(A @ W.T + bias)
    of, oa = _gemm(lib, A.to(cuda), W.to(cuda), bias=bias.to(cuda), gelu=True, prec=1)
    assert metrics(of, ref)["max_rel"] < 1e-5
    assert torch.equal(of, oa)


# ---------------------------------------------------------------- attention
def _attn_ref(qkv, B, S):
    q, k, v = qkv.float().reshape(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
    att = (q @ k.transpose(-1, -2) * 0.125).softmax(-1)
    return (att @ v).permute(0, 2, 1, 3).reshape(B * S, 768)


# 2049 / 257: one scalar tail key + one scalar tail query row; 130 / 2050: two of each; 66: tail keys only;
# 131 / 192 / 1: remainders that take the masked last step and a partial last query tile
# 33 / 97 / 31: sequences whose last step leaves the second key half (keys 32..63 of the step) without a single valid key
# -- the split kernel's half-B warps then carry m = -inf through the merge; 197 = the ViT-B/16 slice trunk (row f-2)
@pytest.mark.parametrize("kernel", ["tri", "split", "rowwarp"])
@pytest.mark.parametrize("B,S,scale", [(2, 2049, 1.0), (1, 2049, 4.0), (3, 128, 2.0), (2, 130, 1.0), (1, 1, 1.0),
                                       (1, 257, 8.0), (2, 66, 2.0), (1, 131, 1.0), (2, 192, 3.0), (1, 2050, 2.0),
                                       (2, 33, 2.0), (1, 97, 1.0), (3, 31, 4.0), (5, 197, 2.0),
                                       # S % 64 == 1: the split kernel folds the orphan key into its epilogue
                                       (2, 129, 2.0), (1, 65, 1.0), (1, 193, 3.0)])
def test_attention_bf16(lib, cuda, B, S, scale, kernel, monkeypatch):
    from hsenet_b200 import _lib
    monkeypatch.setenv("HSENET_ATT_KERNEL", kernel)      # read by the launcher on every call
    g = torch.Generator().manual_seed(S + B)
    qkv = (torch.randn(B * S, 2304, generator=g) * scale).to(torch.bfloat16)
    ref = _attn_ref(qkv, B, S)
    qd = qkv.to(cuda)
    out = torch.full((B * S, 768), float("nan"), dtype=torch.bfloat16, device=cuda)
    _lib.check(lib.hsenet_self_attention(qd.data_ptr(), out.data_ptr(), B, S, 0, _st()), "attention")
    torch.cuda.synchronize()
    m = metrics(out, ref)
    assert m["cos"] > 0.9995 and m["max_rel"] < 1.5e-2, m


# Max-free softmax (scratch given): scale 0.4 keeps c |q| max|k| below the threshold (bounded mode), 1.0 mixes rows, 4.0
# forces the exact fallback everywhere; anti-aligned and zero keys probe the underflow margin of the bound
@pytest.mark.parametrize("B,S,scale", [(2, 2049, 0.4), (2, 2049, 1.0), (1, 2049, 4.0), (3, 197, 0.5), (2, 130, 0.3),
                                       (1, 1, 0.5), (2, 33, 0.6), (1, 2050, 0.45), (2, 129, 0.4)])
def test_attention_bf16_maxfree(lib, cuda, B, S, scale, monkeypatch):
    from hsenet_b200 import _lib
    monkeypatch.setenv("HSENET_ATT_MAXFREE", "1")        # opt-in mode (not faster on B200, kept with its tests)
    g = torch.Generator().manual_seed(7 * S + B)
    qkv = (torch.randn(B * S, 2304, generator=g) * scale)
    if S >= 130:
        qkv[5, 768:1536] = 0.0                                   # a zero key
        qkv[7, 768:768 + 64] = -3.0 * qkv[9, 0:64]               # a key anti-aligned with query 9 of head 0
    qkv = qkv.to(torch.bfloat16)
    ref = _attn_ref(qkv, B, S)
    qd = qkv.to(cuda)
    out = torch.full((B * S, 768), float("nan"), dtype=torch.bfloat16, device=cuda)
    out2 = torch.empty_like(out)
    sp = (S + 127) // 128 * 128
    lse = torch.empty(B, 12, sp, device=cuda)
    scratch = torch.full((B * 12,), float("nan"), device=cuda)
    _lib.check(lib.hsenet_self_attention_ws(qd.data_ptr(), out.data_ptr(), lse.data_ptr(), scratch.data_ptr(), B, S, 0,
                                            _st()), "attention_ws")
    _lib.check(lib.hsenet_self_attention(qd.data_ptr(), out2.data_ptr(), B, S, 0, _st()), "attention")
    torch.cuda.synchronize()
    kn = qkv.float().reshape(B, S, 3, 12, 64)[:, :, 1].square().sum(-1).amax(dim=1).reshape(-1)       # max_k |k|^2 per (b,h)
    assert torch.allclose(scratch.cpu(), kn, rtol=1e-5)
    m = metrics(out, ref)
    assert m["cos"] > 0.9995 and m["max_rel"] < 1.5e-2, m
    assert metrics(out, out2.float().cpu())["max_rel"] < 1.5e-2                # max-free vs exact online softmax
    q, k, _ = qkv.float().reshape(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
    ref_lse = torch.logsumexp(q @ k.transpose(-1, -2) * 0.125, dim=-1) * 1.4426950408889634
    assert torch.allclose(lse[:, :, :S].cpu(), ref_lse, atol=2e-2, rtol=0)


@pytest.mark.parametrize("B,S", [(1, 2049), (2, 130)])
def test_attention_f32(lib, cuda, B, S):
    from hsenet_b200 import _lib
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(B * S, 2304, generator=g) * 2
    ref = _attn_ref(qkv, B, S)
    qd = qkv.to(cuda)
    out = torch.empty(B * S, 768, dtype=torch.float32, device=cuda)
    _lib.check(lib.hsenet_self_attention(qd.data_ptr(), out.data_ptr(), B, S, 1, _st()), "attention")
    assert metrics(out, ref)["max_rel"] < 1e-5


# ---------------------------------------------------------------- row kernels
def test_layernorm(lib, cuda):
    from hsenet_b200 import _lib
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4099, 768, generator=g) * 3 + 1
    gm, bt = torch.randn(768, generator=g), torch.randn(768, generator=g)
    ref = F.layer_norm(x, (768,), gm, bt, 1e-5)
    xd, gd, bd = x.to(cuda), gm.to(cuda), bt.to(cuda)
    o32 = torch.empty(4099, 768, device=cuda)
    _lib.check(lib.hsenet_layernorm(xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), 4099, o32.data_ptr(), 0, _st()), "ln")
    assert metrics(o32, ref)["max_rel"] < 1e-5
    o16 = torch.empty(4099, 768, dtype=torch.bfloat16, device=cuda)
    _lib.check(lib.hsenet_layernorm(xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), 4099, o16.data_ptr(), 1, _st()), "ln")
    assert metrics(o16, ref)["max_rel"] < 5e-3


def test_fold_layernorm(lib, cuda):
    from hsenet_b200 import _lib
    g = torch.Generator().manual_seed(11)
    N, K = 2304, 768
    W = torch.randn(N, K, generator=g) * 0.05
    gm, bt, bias = 1 + 0.2 * torch.randn(K, generator=g), 0.3 * torch.randn(K, generator=g), torch.randn(N, generator=g)
    Wd, gd, bd, biasd = W.to(cuda), gm.to(cuda), bt.to(cuda), bias.to(cuda)
    wf = torch.empty(N, K, dtype=torch.bfloat16, device=cuda)
    cs, bf = torch.empty(N, device=cuda), torch.empty(N, device=cuda)
    for bptr, bref in ((biasd.data_ptr(), bias), (None, torch.zeros(N))):
        _lib.check(lib.hsenet_fold_layernorm(Wd.data_ptr(), gd.data_ptr(), bd.data_ptr(), bptr, N, K, wf.data_ptr(),
                                             cs.data_ptr(), bf.data_ptr(), _st()), "fold")
        want = (W * gm).to(torch.bfloat16)
        assert torch.equal(wf.cpu(), want)                                   # bit-exact rounding of gamma (.) W
        assert metrics(cs, want.float().sum(1))["max_rel"] < 1e-5
        assert metrics(bf, bref + W @ bt)["max_rel"] < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_packer_pool_and_window_attention(lib, cuda, dtype):
    from hsenet_b200 import _lib
    B = 2
    g = torch.Generator().manual_seed(9)
    hr = torch.randn(B, 2048, 768, generator=g).to(dtype)
    ref_lr = O.packer_pool(hr.float())
    hd = hr.to(cuda)
    lr = torch.empty(B, 128, 768, dtype=dtype, device=cuda)
    code = 0 if dtype == torch.float32 else 1
    _lib.check(lib.hsenet_packer_pool(hd.data_ptr(), lr.data_ptr(), B, code, _st()), "pool")
    assert metrics(lr, ref_lr)["max_rel"] < (1e-6 if code == 0 else 5e-3)
    # window attention: q fp32 [B*128,768], kv [B*2048,1536]
    q = torch.randn(B * 128, 768, generator=g)
    kv = torch.randn(B * 2048, 1536, generator=g).to(dtype)
    win = torch.from_numpy(O.packer_window_map().astype(np.int64))
    k = kv.float()[:, :768].reshape(B, 2048, 768)[:, win]
    v = kv.float()[:, 768:].reshape(B, 2048, 768)[:, win]
    ref, _ = O.single_head_attention(q.reshape(B, 128, 1, 768), k, v)
    out = torch.empty(B * 128, 768, dtype=dtype, device=cuda)
    _lib.check(lib.hsenet_packer_window_attention(q.to(cuda).data_ptr(), kv.to(cuda).data_ptr(), out.data_ptr(), B,
                                                  code, _st()), "window_attention")
    assert metrics(out, ref.reshape(B * 128, 768))["max_rel"] < (1e-5 if code == 0 else 5e-3)


def test_slice_cross_attention(lib, cuda):
    from hsenet_b200 import _lib
    B = 2
    g = torch.Generator().manual_seed(2)
    q = torch.randn(B * 2048, 768, generator=g)
    kv = torch.randn(B * 32, 1536, generator=g)
    ref, att = O.single_head_attention(q.reshape(B, 2048, 768), kv[:, :768].reshape(B, 32, 768),
                                       kv[:, 768:].reshape(B, 32, 768))
    out = torch.empty(B * 2048, 768, device=cuda)
    attn = torch.empty(B, 2048, 32, device=cuda)
    _lib.check(lib.hsenet_slice_cross_attention(q.to(cuda).data_ptr(), kv.to(cuda).data_ptr(), out.data_ptr(),
                                                attn.data_ptr(), B, 0, _st()), "xattn")
    assert metrics(out, ref.reshape(B * 2048, 768))["max_rel"] < 1e-5
    assert metrics(attn, att)["max_rel"] < 1e-5


@pytest.mark.parametrize("hw", [(224, 224), (256, 256), (112, 128)])
def test_slice_extract(cuda, hw):
    import hsenet_b200 as H
    x = torch.rand(2, 1, 32, 256, 256, generator=torch.Generator().manual_seed(4))
    ref = O.slice_extract(x, hw)
    got = H.extract_slices(x.to(cuda), hw)
    assert got.shape == ref.shape
    assert_fp32(got, ref, "slice_extract", tol=1e-5)
    gb = H.extract_slices(x.to(cuda), hw, dtype=torch.bfloat16)
    assert metrics(gb, ref)["max_rel"] < 5e-3


def test_gather_rows_strided_view(lib, cuda):
    from hsenet_b200 import _lib
    t = torch.randn(3, 2049, 768, device=cuda)
    v = t[:, 1:]
    out = torch.empty(3, 2048, 768, dtype=torch.bfloat16, device=cuda)
    _lib.check(lib.hsenet_gather_rows(v.data_ptr(), 0, v.stride(0), v.stride(1), 3, 2048, out.data_ptr(), 1, _st()),
               "gather")
    assert torch.equal(out, v.to(torch.bfloat16))
