"""Per-kernel parity on the GPU: every C-ABI entry point (through b200st.kernels.CudaKernels) against the
independent torch restatement in tests/fake_kernels.py, on the same device and inputs.

fp32 tolerance 1e-5 relative L2 (different summation order only); bf16 tolerance 2e-2 (contract)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ks():
    from b200st.kernels import CudaKernels
    from fake_kernels import FakeKernels
    return CudaKernels(), FakeKernels()


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to('cuda').to(dtype)


TOL = {torch.float32: 2e-5, torch.bfloat16: 2e-2}


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('ta,tb', [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (37, 29, 13), (64, 64, 64), (200, 130, 70), (513, 1024, 300),
                                   (3200, 512, 512), (16384, 1024, 80)])
def test_gemm(ks, dtype, ta, tb, M, N, K):
    c, f = ks
    a = rnd(K, M, dtype=dtype) if ta else rnd(M, K, dtype=dtype)
    b = rnd(N, K, dtype=dtype, seed=1) if tb else rnd(K, N, dtype=dtype, seed=1)
    bias = rnd(N, seed=2)
    res = rnd(M, N, dtype=dtype, seed=3)
    y = c.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, residual=res, alpha=0.5)
    yr = f.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, residual=res, alpha=0.5)
    assert rel_err(y, yr) < TOL[dtype]
    y = c.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, relu=True)
    yr = f.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, relu=True)
    assert rel_err(y, yr) < TOL[dtype]


@pytest.mark.parametrize('out_dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('ta,tb', [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize('M,N,K', [(20000, 1000, 328), (32256, 1024, 1024), (19001, 512, 264), (9472, 2048, 256),
                                   (3200, 10000, 512)])
def test_gemm_persistent_tiles(ks, out_dtype, ta, tb, M, N, K):
    """The persistent tcgen05 kernel (>= 4 tiles per SM: 128 x 256 and 128 x 128 tiles, ragged M / N / K edges, every
    transpose form, fused bias / residual / ReLU / ReLU-gate epilogues) against the torch restatement and against the
    one-tile-per-CTA kernel."""
    c, f = ks
    dt = torch.bfloat16
    a = rnd(K, M, dtype=dt) if ta else rnd(M, K, dtype=dt)
    b = rnd(N, K, dtype=dt, seed=1) if tb else rnd(K, N, dtype=dt, seed=1)
    bias = rnd(N, seed=2)
    res = rnd(M, N, dtype=out_dtype, seed=3)
    yr = f.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, residual=res, alpha=0.5, out_dtype=out_dtype)
    y = c.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, residual=res, alpha=0.5, out_dtype=out_dtype)
    assert rel_err(y, yr) < TOL[torch.bfloat16]
    old = c.set_gemm_persistent(0)          # mask: bit 0 persistent kernels, bit 1 CTA-pair kernel; default 3
    try:
        y0 = c.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, residual=res, alpha=0.5, out_dtype=out_dtype)
        c.set_gemm_persistent(1)
        y1 = c.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, residual=res, alpha=0.5, out_dtype=out_dtype)
    finally:
        c.set_gemm_persistent(old)
    assert old == 3
    assert rel_err(y, y0) < 1e-6 and rel_err(y1, y0) < 1e-6        # same k order per output element: (near) identical
    y = c.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, relu=True, out_dtype=out_dtype)
    assert rel_err(y, f.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, relu=True, out_dtype=out_dtype)) < TOL[torch.bfloat16]
    if not ta:
        h = torch.relu(rnd(M, N, dtype=out_dtype, seed=4))
        y = c.gemm(a, b, trans_b=tb, relu_gate=h)
        assert rel_err(y, f.gemm(a, b, trans_b=tb, relu_gate=h)) < TOL[torch.bfloat16]
        assert float(y[h == 0].abs().sum()) == 0.0


@pytest.mark.parametrize('ta,tb', [(True, False), (False, True), (True, True), (False, False)])
@pytest.mark.parametrize('M,N,K', [(1024, 1024, 32256), (1024, 256, 64512), (512, 512, 8200), (1000, 700, 4100)])
def test_gemm_pair_split_k_weight_gradients(ks, ta, tb, M, N, K):
    """Long-K, few-tile products with fp32 output (the weight gradients dW = dG^T X): the CTA-pair kernel with the K
    range split over the SM pairs and an fp32 reduction epilogue, against the torch restatement and the one-tile
    split-K kernel."""
    c, f = ks
    dt = torch.bfloat16
    a = rnd(K, M, dtype=dt, scale=0.1) if ta else rnd(M, K, dtype=dt, scale=0.1)
    b = rnd(N, K, dtype=dt, seed=1) if tb else rnd(K, N, dtype=dt, seed=1)
    y = c.gemm(a, b, trans_a=ta, trans_b=tb, out_dtype=torch.float32)
    yr = f.gemm(a, b, trans_a=ta, trans_b=tb, out_dtype=torch.float32)
    assert rel_err(y, yr) < 2e-3                                   # bf16 products, fp32 sums in a different order
    old = c.set_gemm_persistent(0)
    try:
        y0 = c.gemm(a, b, trans_a=ta, trans_b=tb, out_dtype=torch.float32)
    finally:
        c.set_gemm_persistent(old)
    assert rel_err(y, y0) < 1e-4                                   # fp32 partial sums in a different order
    out = torch.full((M, N), 7.0, device='cuda')                   # a caller-provided output is overwritten, not added to
    c.gemm(a, b, trans_a=ta, trans_b=tb, out=out)
    assert rel_err(out, y) < 1e-4


def test_gemm2_persistent_two_segment(ks):
    c, f = ks
    dt = torch.bfloat16
    M, N, K, K2 = 32256, 1024, 1024, 1024
    a, a2 = rnd(M, K, dtype=dt), rnd(M, K2, dtype=dt, seed=1)
    b, b2 = rnd(K, N, dtype=dt, seed=2), rnd(K2, N, dtype=dt, seed=3)
    assert rel_err(c.gemm2(a, b, a2, b2), f.gemm2(a, b, a2, b2)) < TOL[dt]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,N,K', [(3200, 2048, 512), (37, 29, 13), (64, 512, 2048)])
def test_gemm_relu_gate_epilogue(ks, dtype, M, N, K):
    """relu mode 2: the residual slot carries the forward activation and gates the product (ReLU backward)."""
    c, f = ks
    a, b = rnd(M, K, dtype=dtype), rnd(K, N, dtype=dtype, seed=1)
    h = torch.relu(rnd(M, N, dtype=dtype, seed=2))
    y = c.gemm(a, b, relu_gate=h)
    yr = f.gemm(a, b, relu_gate=h)
    assert rel_err(y, yr) < TOL[dtype]
    assert float(y[h == 0].abs().sum()) == 0.0


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,N,K,K2,tb', [(8064, 1024, 1024, 1024, False), (300, 80, 1024, 1024, False),
                                         (64, 512, 128, 200, True), (37, 29, 13, 64, False), (513, 256, 64, 40, True)])
def test_gemm2_two_segment(ks, dtype, M, N, K, K2, tb):
    """C = A B + A2 B2 (+ bias + residual) in one launch: two-segment K loop on the tensor-core path, chained
    GEMMs otherwise (fp32, first segment not a multiple of 64, ...)."""
    c, f = ks
    a, a2 = rnd(M, K, dtype=dtype), rnd(M, K2, dtype=dtype, seed=1)
    b = rnd(N, K, dtype=dtype, seed=2) if tb else rnd(K, N, dtype=dtype, seed=2)
    b2 = rnd(N, K2, dtype=dtype, seed=3) if tb else rnd(K2, N, dtype=dtype, seed=3)
    res, bias = rnd(M, N, dtype=dtype, seed=4), rnd(N, seed=5)
    assert rel_err(c.gemm2(a, b, a2, b2, trans_b=tb), f.gemm2(a, b, a2, b2, trans_b=tb)) < TOL[dtype]
    y = c.gemm2(a, b, a2, b2, trans_b=tb, bias=bias, residual=res, alpha=0.5)
    yr = f.gemm2(a, b, a2, b2, trans_b=tb, bias=bias, residual=res, alpha=0.5)
    assert rel_err(y, yr) < TOL[dtype]


def test_gemm_strided_views_and_batched(ks):
    c, f = ks
    w = rnd(96, 50)
    x = rnd(33, 20)
    # column-sliced weight (ldb > K), output into a column slice (ldc > N), in-place accumulate
    out = torch.zeros(33, 128, device='cuda')
    c.gemm(x, w[:, 30:], trans_b=True, out=out[:, 16:112])
    c.gemm(x, w[:, 10:30], trans_b=True, residual=out[:, 16:112], out=out[:, 16:112])
    ref = x @ w[:, 30:].t() + x @ w[:, 10:30].t()
    assert rel_err(out[:, 16:112], ref) < 2e-5 and float(out[:, :16].abs().sum()) == 0
    # batched with permuted (strided) operands, transposed A: [B][S,T]^T @ [B][S,D]
    a = rnd(7, 5, 11).permute(1, 0, 2)
    b = rnd(7, 5, 13, seed=4).permute(1, 0, 2)
    y = c.gemm(a, b, trans_a=True)
    assert rel_err(y, torch.matmul(a.transpose(1, 2), b)) < 2e-5
    # bf16 operands, fp32 output (weight gradients)
    a16, b16 = rnd(300, 40, dtype=torch.bfloat16), rnd(300, 24, dtype=torch.bfloat16, seed=5)
    y = c.gemm(a16, b16, trans_a=True, out_dtype=torch.float32)
    assert y.dtype == torch.float32 and rel_err(y, a16.float().t() @ b16.float()) < 1e-5


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('rows,cols', [(1, 8), (50, 32), (3200, 512), (77, 1000)])
def test_layernorm(ks, dtype, rows, cols):
    c, f = ks
    x = rnd(rows, cols, dtype=dtype, scale=2.0) + 0.5
    g, b = rnd(cols, seed=1) + 1, rnd(cols, seed=2)
    y, mean, rstd = c.layernorm_fwd(x, g, b, 1e-6)
    yr, meanr, rstdr = f.layernorm_fwd(x, g, b, 1e-6)
    assert rel_err(y, yr) < TOL[dtype] and rel_err(mean, meanr) < 1e-5 and rel_err(rstd, rstdr) < 1e-5
    dy = rnd(rows, cols, dtype=dtype, seed=3)
    dg, db = torch.zeros(cols, device='cuda'), torch.zeros(cols, device='cuda')
    dgr, dbr = torch.zeros(cols, device='cuda'), torch.zeros(cols, device='cuda')
    dx = c.layernorm_bwd(dy, x, g, mean, rstd, dg, db)
    dxr = f.layernorm_bwd(dy, x, g, meanr, rstdr, dgr, dbr)
    assert rel_err(dx, dxr) < TOL[dtype] and rel_err(dg, dgr) < 1e-4 and rel_err(db, dbr) < 1e-4


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('B,H,Lq,Lk,d,mask_kind', [(2, 4, 7, 7, 8, 'causal'), (3, 8, 50, 31, 64, 'key'),
                                                   (64, 8, 50, 50, 64, 'causal'), (2, 8, 9, 12, 6, None),
                                                   (2, 2, 5, 150, 64, 'key'), (2, 8, 32, 32, 64, None),
                                                   (2, 8, 64, 64, 64, 'causal'), (3, 8, 33, 49, 64, 'key'),
                                                   (5, 8, 1, 17, 64, 'key')])
def test_mha(ks, dtype, B, H, Lq, Lk, d, mask_kind):
    c, f = ks
    qkv = rnd(B, max(Lq, Lk), 3 * H * d, dtype=dtype)            # packed buffer: exercises row strides
    q, k, v = qkv[:, :Lq, :H * d], qkv[:, :Lk, H * d:2 * H * d], qkv[:, :Lk, 2 * H * d:]
    if Lq != Lk:
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    mask = None
    if mask_kind == 'key':
        lens = torch.randint(1, Lk + 1, (B,), device='cuda')
        mask = (torch.arange(Lk, device='cuda')[None, :] < lens[:, None]).unsqueeze(1).to(torch.uint8)
    elif mask_kind == 'causal':
        ids = torch.randint(0, 3, (B, Lk), device='cuda'); ids[:, 0] = 2
        mask = f.token_mask(ids, 0, True)
    temp = d ** 0.5
    o, p = c.mha_fwd(q, k, v, mask, H, temp)
    orf, pr = f.mha_fwd(q, k, v, mask, H, temp)
    assert rel_err(o, orf) < TOL[dtype] and rel_err(p, pr) < TOL[dtype]
    do = rnd(B, Lq, H * d, dtype=dtype, seed=9)
    dq, dk, dv = c.mha_bwd(do, q, k, v, pr, H, temp)
    dqr, dkr, dvr = f.mha_bwd(do, q, k, v, pr, H, temp)
    for a, b in ((dq, dqr), (dk, dkr), (dv, dvr)):
        assert rel_err(a, b) < TOL[dtype]


@pytest.mark.parametrize('Lq,Lk,mask_kind', [(50, 50, 'causal'), (50, 32, 'key'), (32, 32, 'key')])
def test_mha_tensor_core_matches_cuda_core_tiles(ks, Lq, Lk, mask_kind):
    """The tcgen05 attention core (bf16, d=64, L<=64) against the CUDA-core tile kernels on the training shapes."""
    c, f = ks
    B, H, d = 16, 8, 64
    q, k, v = rnd(B, Lq, H * d, dtype=torch.bfloat16), rnd(B, Lk, H * d, dtype=torch.bfloat16, seed=1), rnd(B, Lk, H * d, dtype=torch.bfloat16, seed=2)
    if mask_kind == 'key':
        lens = torch.randint(1, Lk + 1, (B,), device='cuda')
        mask = (torch.arange(Lk, device='cuda')[None, :] < lens[:, None]).unsqueeze(1).to(torch.uint8)
    else:
        ids = torch.randint(0, 3, (B, Lk), device='cuda'); ids[:, 0] = 2
        mask = f.token_mask(ids, 0, True)
    do = rnd(B, Lq, H * d, dtype=torch.bfloat16, seed=9)
    res = []
    for mode in (0, 1):
        old = c.set_mha_backend(mode)
        try:
            o, p = c.mha_fwd(q, k, v, mask, H, 8.0)
            res.append((o, p) + tuple(c.mha_bwd(do, q, k, v, p, H, 8.0)))
        finally:
            c.set_mha_backend(old)
    for a, b in zip(*res):
        assert rel_err(a, b) < 1e-2


def test_mha_fully_masked_row_is_uniform(ks):
    """-1e9 fill is finite: a row with every key masked softmaxes to uniform (layers.py:224)."""
    c, f = ks
    q, k, v = rnd(1, 3, 16), rnd(1, 5, 16, seed=1), rnd(1, 5, 16, seed=2)
    mask = torch.zeros(1, 1, 5, dtype=torch.uint8, device='cuda')
    o, p = c.mha_fwd(q, k, v, mask, 2, 2.0)
    assert torch.allclose(p, torch.full_like(p, 0.2), atol=1e-6)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_lstm_cell(ks, dtype):
    c, f = ks
    B, H = 5, 24
    gates, cp, res = rnd(B, 4 * H, dtype=dtype), rnd(B, H, seed=1), rnd(B, H, dtype=dtype, seed=2)
    h, cc, acts, outr = c.lstm_cell_fwd(gates, cp, residual=res)
    hr, ccr, actsr, outrr = f.lstm_cell_fwd(gates, cp, residual=res)
    for a, b in ((h, hr), (cc, ccr), (acts, actsr), (outr, outrr)):
        assert rel_err(a, b) < TOL[dtype]
    gb, gc = rnd(B, 4 * H, dtype=dtype, seed=7), rnd(B, 4 * H, dtype=dtype, seed=8)
    h3, c3, _, _ = c.lstm_cell_fwd(gates, cp, gates_b=gb, gates_c=gc)
    hr3, cr3, _, _ = f.lstm_cell_fwd(gates, cp, gates_b=gb, gates_c=gc)
    assert rel_err(h3, hr3) < TOL[dtype] and rel_err(c3, cr3) < TOL[dtype]
    h0, c0, _, _ = c.lstm_cell_fwd(gates, None)
    hr0, cr0, _, _ = f.lstm_cell_fwd(gates, None)
    assert rel_err(h0, hr0) < TOL[dtype]
    dhs = [rnd(B, H, dtype=dtype, seed=3), None, rnd(B, H, dtype=dtype, seed=4)]
    dcn = rnd(B, H, seed=5)
    dg, dcp = c.lstm_cell_bwd(dhs, dcn, actsr, cp, ccr, dtype)
    dgr, dcpr = f.lstm_cell_bwd(dhs, dcn, actsr, cp, ccr, dtype)
    assert rel_err(dg, dgr) < TOL[dtype] and rel_err(dcp, dcpr) < 1e-5


@pytest.fixture(params=[0, 3], ids=['default', 'warp-mma'])
def blstm_backend(request, ks):
    """0 = the default kernels (tcgen05 for bf16 / H = 256, CUDA cores otherwise), 3 = the register-resident warp-MMA
    kernels of csrc/lstm_rg.cu (bf16 / H = 256; other shapes fall through to the same kernels as 0)."""
    old = ks[0].set_blstm_backend(request.param)
    yield request.param
    ks[0].set_blstm_backend(old)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('T,B,H,pair', [(8, 3, 16, 2), (24, 5, 24, 2), (6, 11, 16, 1), (40, 9, 32, 2),
                                        (16, 8, 256, 2), (64, 40, 256, 2), (24, 16, 256, 1), (2, 1, 256, 2), (1, 19, 256, 1)])
def test_blstm_recurrence(ks, blstm_backend, dtype, T, B, H, pair):
    c, f = ks
    if blstm_backend == 3 and not (dtype == torch.bfloat16 and H == 256):
        pytest.skip('the warp-MMA kernels take bf16 / H = 256 only')
    xproj = rnd(2, T, B, 4 * H, dtype=dtype)
    wf, wr = rnd(4 * H, H, seed=1, scale=H ** -0.5), rnd(4 * H, H, seed=2, scale=H ** -0.5)
    lens = torch.randint(1, T + 1, (B,), device='cuda', dtype=torch.int32)
    lens[0] = T
    if pair == 2:
        shape, ld_t, ld_b = (T // 2, B, 4 * H), B * 4 * H, 4 * H
    else:
        shape, ld_t, ld_b = (B, T, 2 * H), 2 * H, T * 2 * H
    out = torch.full(shape, 7.0, dtype=dtype, device='cuda')
    outr = torch.full(shape, 7.0, dtype=dtype, device='cuda')
    hs, acts, cs = c.blstm_fwd(xproj, wf, wr, lens, out, ld_t, ld_b, pair)
    hsr, actsr, csr = f.blstm_fwd(xproj, wf, wr, lens, outr, ld_t, ld_b, pair)
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    assert rel_err(out, outr) < tol
    assert rel_err(hs, hsr) < tol
    valid = (torch.arange(T, device='cuda')[:, None] < lens[None, :]).float()[None, :, :, None]
    blocked = c.blstm_saved_blocked(dtype, H)       # the saved state is kernel-private: compare / feed it through the documented layout
    if blocked:
        acts, cs = c.blstm_unblock(acts, cs, B)
    assert rel_err(acts * valid, actsr * valid) < tol and rel_err(cs * valid, csr * valid) < tol
    dout = rnd(*shape, dtype=dtype, seed=7)
    a_in, c_in = c.blstm_block(actsr, csr) if blocked else (actsr, csr)
    if blocked:
        ra, rc = c.blstm_unblock(a_in, c_in, B)
        assert torch.equal(ra, actsr) and torch.equal(rc, csr)
    dg = c.blstm_bwd(dout, ld_t, ld_b, pair, a_in, c_in, wf, wr, lens, dtype)
    dgr = f.blstm_bwd(dout, ld_t, ld_b, pair, actsr, csr, wf, wr, lens, dtype)
    assert rel_err(dg, dgr) < tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_las_attention(ks, dtype):
    c, f = ks
    B, Tk, D, Dv = 6, 13, 32, 48
    q, wk, vals = rnd(B, D, dtype=dtype), rnd(B, Tk, D, dtype=dtype, seed=1), rnd(B, Tk, Dv, dtype=dtype, seed=2)
    klens = torch.tensor([13, 1, 5, 7, 13, 2], dtype=torch.int32, device='cuda')
    cx, p = c.las_attn_fwd(q, wk, vals, klens)
    cxr, pr = f.las_attn_fwd(q, wk, vals, klens)
    assert rel_err(cx, cxr) < TOL[dtype] and rel_err(p, pr) < TOL[dtype]
    assert float(p[1, 1:].abs().sum()) == 0.0
    dctx = rnd(B, Dv, dtype=dtype, seed=3)
    ds, dq = c.las_attn_bwd(dctx, wk, vals, pr)
    dsr, dqr = f.las_attn_bwd(dctx, wk, vals, pr)
    assert rel_err(ds, dsr) < TOL[dtype] and rel_err(dq, dqr) < TOL[dtype]


@pytest.mark.parametrize('B,Tk,D,Dv', [(64, 126, 512, 512), (5, 37, 64, 96), (3, 4, 32, 32), (2, 256, 512, 512)])
def test_las_attention_key_split_cluster(ks, B, Tk, D, Dv):
    """bf16 LAS attention step on the 4-CTA key-split cluster kernels (training shape first) against the torch
    restatement and against the one-CTA-per-sequence kernels; includes klen = 0 (uniform row) and klen = Tk."""
    c, f = ks
    dt = torch.bfloat16
    q, wk, vals = rnd(B, D, dtype=dt), rnd(B, Tk, D, dtype=dt, seed=1), rnd(B, Tk, Dv, dtype=dt, seed=2)
    klens = torch.randint(1, Tk + 1, (B,), device='cuda').to(torch.int32)
    klens[0] = Tk
    klens[-1] = 0
    dctx = rnd(B, Dv, dtype=dt, seed=3)
    res = []
    for mode in (0, 2):
        old = c.set_mha_backend(mode)
        try:
            cx, p = c.las_attn_fwd(q, wk, vals, klens)
            ds, dq = c.las_attn_bwd(dctx, wk, vals, p)
            res.append((cx, p, ds, dq))
        finally:
            c.set_mha_backend(old)
    cxr, pr = f.las_attn_fwd(q, wk, vals, klens)
    dsr, dqr = f.las_attn_bwd(dctx, wk, vals, res[0][1])
    for a, b in zip(res[0], (cxr, pr, dsr, dqr)):
        assert rel_err(a, b) < TOL[dt]
    for a, b in zip(*res):
        assert rel_err(a, b) < TOL[dt]
    assert abs(float(res[0][1][-1].sum()) - 1.0) < 1e-3 and float((res[0][1][-1] - 1.0 / Tk).abs().max()) < 1e-6


def test_argmax_and_lengths(ks):
    c, f = ks
    x = rnd(9, 1000)
    x[3, 17] = 50.0; x[3, 400] = 50.0            # tie: first index wins
    idx = torch.empty(4, 9, dtype=torch.int64, device='cuda')
    c.argmax_rows(x, idx[2])
    assert torch.equal(idx[2], x.argmax(1)) and int(idx[2, 3]) == 17
    col = torch.empty(9, 3, dtype=torch.int64, device='cuda')
    c.argmax_rows(x, col[:, 1])
    assert torch.equal(col[:, 1], x.argmax(1))
    big = rnd(5, 10000, dtype=torch.bfloat16)
    big[0, 3] = 100.0; big[2, 0] = 100.0; big[4, 3] = 100.0; big[3, 9999] = 100.0
    ln = torch.tensor([7, 7, 7, 7, 2], dtype=torch.int32, device='cuda')
    out = torch.empty(5, dtype=torch.int64, device='cuda')
    c.argmax_rows(big, out, lengths=ln, step=3)
    assert out.tolist() == [3, int(big[1].float().argmax()), 0, 9999, 3] and ln.tolist() == [4, 7, 4, 7, 2]
    table = rnd(10000, 200)
    emb = torch.empty(5, 200, dtype=torch.bfloat16, device='cuda')
    out2 = torch.empty(5, dtype=torch.int64, device='cuda')
    c.argmax_rows(big, out2, embed=(table, emb))                  # arg-max + embedding feed in one launch
    assert torch.equal(out2, out) and torch.equal(emb, table[out].to(torch.bfloat16))
    sym = torch.tensor([3, 5, 0, 9, 3], device='cuda')
    lengths = torch.tensor([7, 7, 7, 7, 2], dtype=torch.int32, device='cuda')
    c.las_update_lengths(sym, lengths, 3)
    assert lengths.tolist() == [4, 7, 4, 7, 2]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_embedding_and_mix(ks, dtype):
    c, f = ks
    V, E, D, n = 50, 12, 20, 37
    table = rnd(V, E)
    ids = torch.randint(0, V, (n,), device='cuda'); ids[:5] = 0
    assert rel_err(c.embedding_fwd(ids, table, dtype), f.embedding_fwd(ids, table, dtype)) < 1e-6 + TOL[dtype]
    buf = torch.zeros(n, 40, dtype=dtype, device='cuda')
    c.embedding_fwd(ids, table, dtype, out=buf[:, 4:4 + E])
    assert rel_err(buf[:, 4:4 + E], table[ids].to(dtype)) < 1e-6 and float(buf[:, :4].abs().sum()) == 0
    dout = rnd(n, E, dtype=dtype, seed=1)
    dt1, dt2 = torch.zeros(V, E, device='cuda'), torch.zeros(V, E, device='cuda')
    c.embedding_bwd(ids, dout, dt1, 0)
    f.embedding_bwd(ids, dout, dt2, 0)
    assert rel_err(dt1, dt2) < 1e-5 and float(dt1[0].abs().sum()) == 0.0
    dyn = rnd(n, D, dtype=dtype, seed=2)
    assert rel_err(c.mix_gather_concat(ids, table, dyn), f.mix_gather_concat(ids, table, dyn)) < 1e-6 + TOL[dtype]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('rows,cols', [(5, 41), (64, 10000), (3, 33000), (7, 2048), (9, 6000), (3136, 10000)])
def test_softmax_family(ks, dtype, rows, cols):
    c, f = ks
    x = rnd(rows, cols, dtype=dtype, scale=3.0)
    y, am = c.log_softmax_fwd(x, want_argmax=True)
    yr, amr = f.log_softmax_fwd(x, want_argmax=True)
    assert rel_err(y, yr) < TOL[dtype]
    if dtype == torch.float32:
        assert torch.equal(am, amr)
    dy = rnd(rows, cols, dtype=dtype, seed=1)
    assert rel_err(c.log_softmax_bwd(dy, yr), f.log_softmax_bwd(dy, yr)) < TOL[dtype]
    tgt = torch.randint(0, cols, (rows,), device='cuda')
    mask = (torch.arange(rows, device='cuda') % 3 != 1).to(torch.uint8)
    l, lr = c.masked_nll_fwd(yr, tgt, mask), f.masked_nll_fwd(yr, tgt, mask)
    assert rel_err(l, lr) < 1e-5
    g = torch.tensor([0.25], device='cuda')
    assert rel_err(c.masked_nll_bwd(g, tgt, mask, rows, cols, dtype), f.masked_nll_bwd(g, tgt, mask, rows, cols, dtype)) < 1e-6
    for eps in (0.0, 0.1):
        ls, d = c.softmax_nll_fused(x, tgt, mask, g, eps)
        lsr, dr = f.softmax_nll_fused(x, tgt, mask, g, eps)
        assert rel_err(ls, lsr) < 1e-5 and rel_err(d, dr) < TOL[dtype]
    # fused == log_softmax + nll  (eps = 0)
    ls, _ = c.softmax_nll_fused(x, tgt, mask, g, 0.0)
    assert rel_err(ls, lr) < (1e-2 if dtype == torch.bfloat16 else 1e-5)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_glue(ks, dtype):
    c, f = ks
    a, b = rnd(7, 33, dtype=dtype), rnd(7, 33, dtype=dtype, seed=1)
    assert rel_err(c.add(a, b), f.add(a, b)) < 1e-6 + TOL[dtype] * 0.5
    x, pe = rnd(3, 9, 16, dtype=dtype), rnd(500, 16, seed=2)
    assert rel_err(c.add_posenc(x, pe), f.add_posenc(x, pe)) < 1e-6 + TOL[dtype] * 0.5
    t = rnd(5, 7, 11, dtype=dtype)
    assert torch.equal(c.transpose01(t), f.transpose01(t))
    assert torch.equal(c.transpose01(t.float(), out_dtype=dtype), f.transpose01(t.float(), out_dtype=dtype))
    assert torch.equal(c.cast(t.float(), dtype), t.float().to(dtype))
    big = rnd(5000, 130, dtype=dtype)
    assert rel_err(c.colsum(big), f.colsum(big)) < 1e-4
    acc = torch.ones(130, device='cuda')
    c.colsum(big, out=acc, accumulate=True)
    assert rel_err(acc, f.colsum(big) + 1) < 1e-4
    for r_, c_ in ((3200, 2048), (77, 520), (1, 8), (4100, 1000)):       # 16-byte-vector path (bf16) + tails
        wide = rnd(r_, c_ + 8, dtype=dtype)[:, :c_]                       # ld > cols
        assert rel_err(c.colsum(wide), f.colsum(wide)) < 1e-4
    assert torch.equal(c.relu_bwd(a, b), f.relu_bwd(a, b))
    ids = torch.randint(0, 4, (6, 9), device='cuda')
    assert torch.equal(c.token_mask(ids, 0, True), f.token_mask(ids, 0, True))
    assert torch.equal(c.token_mask(ids, 0, False), f.token_mask(ids, 0, False))
    ln = torch.tensor([1, 9, 4], dtype=torch.int32, device='cuda')
    assert torch.equal(c.length_mask(ln, 9), f.length_mask(ln, 9))


@pytest.mark.parametrize('ta,tb', [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (100, 72, 80), (3200, 512, 512), (64, 2048, 712),
                                   (64, 10000, 512), (1024, 80, 4096), (1024, 1024, 16384), (37, 24, 8),
                                   # decoder-step shapes served by the cluster split-K kernel (M <= 64, K >= 512)
                                   (64, 512, 2048), (64, 1024, 2048), (37, 200, 1000), (64, 512, 512), (5, 72, 2048),
                                   (64, 2048, 1536),
                                   # weight-gradient shapes (ta, !tb, fp32 out): cluster split-K with 128-row tiles
                                   (512, 512, 3200), (2048, 512, 2048), (512, 2048, 3200), (1024, 256, 8064)])
def test_gemm_tensor_core(ks, ta, tb, M, N, K):
    """tcgen05/TMA/TMEM kernel (forced) vs fp32 matmul of the same bf16 operands: K-major and MN-major
    operand staging, M/N/K tails (TMA zero fill), bf16 and fp32 (split-K atomics) outputs."""
    c, f = ks
    a = rnd(K, M, dtype=torch.bfloat16) if ta else rnd(M, K, dtype=torch.bfloat16)
    b = rnd(N, K, dtype=torch.bfloat16, seed=1) if tb else rnd(K, N, dtype=torch.bfloat16, seed=1)
    if a.stride(0) % 8 or b.stride(0) % 8:
        pytest.skip('leading dimension not TMA addressable: served by the CUDA-core kernel')
    old = c.set_gemm_backend(2)
    try:
        y16 = c.gemm(a, b, trans_a=ta, trans_b=tb)
        y32 = c.gemm(a, b, trans_a=ta, trans_b=tb, out_dtype=torch.float32)
        bias, res = rnd(N, seed=2), rnd(M, N, dtype=torch.bfloat16, seed=3)
        ye = c.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, residual=res, alpha=0.5)
        yr = c.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, relu=True)
    finally:
        c.set_gemm_backend(old)
    ref = (a.float().t() if ta else a.float()) @ (b.float().t() if tb else b.float())
    assert rel_err(y32, ref) < 1e-4
    assert rel_err(y16, ref) < 1e-2
    assert rel_err(ye, 0.5 * ref + bias + res.float()) < 1e-2
    assert rel_err(yr, torch.relu(ref + bias)) < 1e-2


def test_gemm_tensor_core_views(ks):
    c, f = ks
    a, w = rnd(300, 712, dtype=torch.bfloat16), rnd(512, 712, dtype=torch.bfloat16, seed=1)
    old = c.set_gemm_backend(2)
    try:
        y = c.gemm(a[:, 200:], w[:, 200:], trans_b=True)
        out = torch.zeros(300, 1024, device='cuda', dtype=torch.bfloat16)
        c.gemm(a, w, trans_b=True, out=out[:, 256:768])
        c.gemm(a, w, trans_b=True, residual=out[:, 256:768], out=out[:, 256:768])
        with pytest.raises(RuntimeError):
            c.gemm(a[:, 3:], w[:, 3:], trans_b=True)        # misaligned base: not TMA addressable
    finally:
        c.set_gemm_backend(old)
    assert rel_err(y, a[:, 200:].float() @ w[:, 200:].float().t()) < 1e-2
    assert rel_err(out[:, 256:768], 2 * (a.float() @ w.float().t())) < 1e-2
    assert float(out[:, :256].abs().sum()) == 0 and float(out[:, 768:].abs().sum()) == 0


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('n_hyp,bdiv,H,d,Lk,Lmax', [(6, 1, 4, 16, 5, 9), (15, 5, 8, 64, 31, 31), (640, 5, 8, 64, 49, 50),
                                                    (4, 1, 2, 8, 1, 3)])
def test_mha_decode_cached_attention(ks, dtype, n_hyp, bdiv, H, d, Lk, Lmax):
    """Single-query attention over a K|V cache: self-attention form (ancestry table, PAD-key mask) and cross-attention
    form (beams share the utterance's keys: slot = b // bdiv, mask row b // bdiv)."""
    c, f = ks
    HD = H * d
    g = torch.Generator().manual_seed(n_hyp + Lk)
    q = torch.randn(n_hyp, HD, generator=g).cuda().to(dtype)
    # self-attention: one [n_hyp, Lmax, 2*HD] cache, K and V are its column halves
    cache = torch.randn(n_hyp, Lmax, 2 * HD, generator=g).cuda().to(dtype)
    anc = torch.stack([torch.randperm(n_hyp, generator=g) for _ in range(Lmax)]).to(torch.int32).cuda()
    mask = (torch.rand(n_hyp, Lmax, generator=g) > 0.3).to(torch.uint8).cuda()
    mask[:, 0] = 1
    for a, m in ((anc, mask), (None, None), (anc, None)):
        got = c.mha_decode(q, cache[:, :, :HD], cache[:, :, HD:], Lk, H, d ** 0.5, anc=a, mask=m)
        ref = f.mha_decode(q, cache[:, :, :HD], cache[:, :, HD:], Lk, H, d ** 0.5, anc=a, mask=m)
        assert rel_err(got, ref) < TOL[dtype]
    # cross-attention: n_hyp // bdiv utterances
    if n_hyp % bdiv == 0:
        nb = n_hyp // bdiv
        kv = torch.randn(nb, Lk, 2 * HD, generator=g).cuda().to(dtype)
        smask = (torch.rand(nb, 1, Lk, generator=g) > 0.3).to(torch.uint8).cuda()
        smask[:, :, 0] = 1
        got = c.mha_decode(q, kv[:, :, :HD], kv[:, :, HD:], Lk, H, d ** 0.5, bdiv=bdiv, mask=smask, mask_bdiv=bdiv)
        ref = f.mha_decode(q, kv[:, :, :HD], kv[:, :, HD:], Lk, H, d ** 0.5, bdiv=bdiv, mask=smask, mask_bdiv=bdiv)
        assert rel_err(got, ref) < TOL[dtype]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('rows,cols', [(3200, 512), (37, 128), (1984, 1024)])
def test_layernorm_bwd_partial_sums(ks, dtype, rows, cols):
    """LayerNorm backward with per-CTA dgamma / dbeta partial sums (column-summed by the caller) == the atomics version."""
    c, f = ks
    x, dy, add = rnd(rows, cols, dtype=dtype), rnd(rows, cols, dtype=dtype, seed=1), rnd(rows, cols, dtype=dtype, seed=2)
    gamma = rnd(cols, seed=3)
    _, mean, rstd = c.layernorm_fwd(x, gamma, rnd(cols, seed=4), 1e-6)
    dx, part = c.layernorm_bwd_partial(dy, x, gamma, mean, rstd, add=add)
    dg, db = torch.zeros(cols, device='cuda'), torch.zeros(cols, device='cuda')
    dx_ref = f.layernorm_bwd(dy, x, gamma, mean, rstd, dg, db, add=add)
    assert part.shape == ((rows + 15) // 16, 2 * cols)
    assert rel_err(dx, dx_ref) < TOL[dtype]
    assert rel_err(c.colsum(part[:, :cols]), dg) < 1e-4 and rel_err(c.colsum(part[:, cols:]), db) < 1e-4
    assert c.layernorm_bwd_partial(dy[:, :100].contiguous(), x[:, :100].contiguous(), gamma[:100], mean, rstd) == (None, None)


@pytest.mark.parametrize('M,K,with_bias,with_res', [(3200, 512, False, True), (2048, 1024, True, True), (50, 512, True, False),
                                                   (129, 72, False, True), (1, 64, True, True), (300, 1024, False, False)])
def test_gemm_ln_fused_epilogue(ks, M, K, with_bias, with_res):
    """b200st_gemm_ln == b200st_gemm(+bias, +residual) followed by b200st_layernorm_fwd (layers.py:190-197,245-252 + 153)."""
    c, f = ks
    a = rnd(M, K, dtype=torch.bfloat16, seed=1)
    w = rnd(512, K, dtype=torch.bfloat16, seed=2, scale=K ** -0.5)
    bias = rnd(512, seed=3) if with_bias else None
    res = rnd(M, 512, dtype=torch.bfloat16, seed=4, scale=2.0) if with_res else None
    gamma, beta = 1 + 0.2 * rnd(512, seed=5), 0.1 * rnd(512, seed=6)
    assert c.gemm_ln_ok(a, w, res)
    y, yn, mean, rstd = c.gemm_ln(a, w, bias, res, gamma, beta, 1e-6)
    y0 = c.gemm(a, w, trans_b=True, bias=bias, residual=res)
    yn0, mean0, rstd0 = c.layernorm_fwd(y0, gamma, beta, 1e-6)
    assert rel_err(y, y0) < 1e-2                       # same fp32 accumulator, bf16 rounding of the sum may differ in the last bit
    # the statistics are those of the fused kernel's OWN rounded y: compare against the LayerNorm of exactly that tensor
    yn1, mean1, rstd1 = c.layernorm_fwd(y, gamma, beta, 1e-6)
    assert rel_err(mean, mean1) < 1e-5 and rel_err(rstd, rstd1) < 1e-5
    assert rel_err(yn, yn1) < 4e-3
    ynr, meanr, rstdr = f.layernorm_fwd(y, gamma, beta, 1e-6)
    assert rel_err(yn, ynr) < 4e-3 and rel_err(mean, meanr) < 1e-5 and rel_err(rstd, rstdr) < 1e-5
    assert rel_err(yn, yn0) < 2e-2


@pytest.mark.parametrize('M,K,with_add', [(3200, 512, True), (2048, 1024, True), (50, 512, False), (129, 72, True), (1, 64, True)])
def test_gemm_lnbwd_fused_epilogue(ks, M, K, with_add):
    """b200st_gemm_lnbwd == b200st_gemm followed by the LayerNorm backward (+ skip gradient) and its dgamma / dbeta sums."""
    c, f = ks
    a = rnd(M, K, dtype=torch.bfloat16, seed=1)
    w = rnd(K, 512, dtype=torch.bfloat16, seed=2, scale=K ** -0.5)
    x = rnd(M, 512, dtype=torch.bfloat16, seed=3, scale=2.0)
    add = rnd(M, 512, dtype=torch.bfloat16, seed=4) if with_add else None
    gamma, beta = 1 + 0.2 * rnd(512, seed=5), 0.1 * rnd(512, seed=6)
    _, mean, rstd = c.layernorm_fwd(x, gamma, beta, 1e-6)
    assert c.gemm_lnbwd_ok(a, w, x, add)
    dx, part = c.gemm_lnbwd(a, w, x, gamma, mean, rstd, add=add)
    dxr, partr = f.gemm_lnbwd(a, w, x, gamma, mean, rstd, add=add)
    assert part.shape == ((M + 127) // 128, 1024)
    assert rel_err(dx, dxr) < 1e-2
    assert rel_err(part.sum(0), partr.sum(0)) < 2e-3
    # and against the two separate kernels it replaces (their dy is rounded to bf16 in between)
    dy = c.gemm(a, w)
    dln = torch.zeros(2, 512, device='cuda')
    dx0 = c.layernorm_bwd(dy, x, gamma, mean, rstd, dln[0], dln[1], add=add)
    assert rel_err(dx, dx0) < 2e-2 and rel_err(part.sum(0), dln.reshape(-1)) < 1e-2


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('S,B,Tk,D', [(31, 64, 126, 512), (5, 3, 13, 48), (70, 2, 17, 33), (1, 1, 1, 8)])
def test_las_stack_grad(ks, dtype, S, B, Tk, D):
    c, f = ks
    w = rnd(S, B, Tk, seed=1)
    x = rnd(S, B, D, dtype=dtype, seed=2)
    assert rel_err(c.las_stack_grad(w, x), f.las_stack_grad(w, x)) < TOL[dtype]


@pytest.mark.parametrize('M,K,with_bias', [(3200, 512, False), (2048, 1024, True), (77, 512, True)])
def test_gemm_ln_fused_dropout(ks, M, K, with_bias):
    """Dropout on the projection output inside the fused epilogue draws the SAME mask as b200st_dropout (layers.py:194-195,248-250)."""
    c, f = ks
    a = rnd(M, K, dtype=torch.bfloat16, seed=1)
    w = rnd(512, K, dtype=torch.bfloat16, seed=2, scale=K ** -0.5)
    bias = rnd(512, seed=3) if with_bias else None
    res = rnd(M, 512, dtype=torch.bfloat16, seed=4, scale=2.0)
    gamma, beta = 1 + 0.2 * rnd(512, seed=5), 0.1 * rnd(512, seed=6)
    rng = torch.tensor([0x1234567, 9], dtype=torch.int64, device='cuda')
    p, site = 0.2, 17
    y, yn, mean, rstd = c.gemm_ln(a, w, bias, res, gamma, beta, 1e-6, dropout=(p, rng, site))
    g0 = c.gemm(a, w, trans_b=True, bias=bias)
    y0 = c.dropout(g0, p, rng, site, residual=res)
    keep = c.dropout(torch.ones_like(g0), p, rng, site) != 0
    # the same elements are dropped (y == residual exactly there) and the kept ones agree to bf16 rounding
    assert torch.equal(y[~keep], res[~keep])
    assert 0.15 < float((~keep).float().mean()) < 0.25
    assert rel_err(y, y0) < 1e-2
    yn1, mean1, rstd1 = c.layernorm_fwd(y, gamma, beta, 1e-6)
    assert rel_err(yn, yn1) < 4e-3 and rel_err(mean, mean1) < 1e-5 and rel_err(rstd, rstd1) < 1e-5
    yr, ynr, _, _ = f.gemm_ln(a, w, bias, res, gamma, beta, 1e-6, dropout=(p, rng, site))
    assert rel_err(y, yr) < 1e-2 and rel_err(yn, ynr) < 2e-2
