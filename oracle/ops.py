"""Functional numpy restatement of the Torch7 ops on the hot path.

Every function works in the dtype of its inputs (float32 = faithful to
``torch.setdefaulttensortype('torch.FloatTensor')`` at ``train.lua:48``; float64
for tight cross-checks).  Layout is the reference's: NCHW, contiguous.
Test infrastructure only -- see ``oracle/__init__.py``.
"""
import numpy as np

def bf16_round(x):
    """Round to the nearest bfloat16 (ties to even), returned in the dtype of x."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32)
    r = ((u + (((u >> 16) & 1) + 0x7FFF)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(a.shape).astype(np.asarray(x).dtype)


# --------------------------------------------------------------------------
# im2col helpers (THNN SpatialConvolutionMM = unfold + sgemm)
# --------------------------------------------------------------------------


def conv_out_size(h, k, d, p):
    """nn.SpatialConvolution: oH = floor((H + 2p - k)/d) + 1."""
    return (h + 2 * p - k) // d + 1


def fullconv_out_size(h, k, d, p, adj=0):
    """nn.SpatialFullConvolution: oH = (H-1)d - 2p + k + adj."""
    return (h - 1) * d - 2 * p + k + adj


def im2col(x, kH, kW, dH, dW, pH, pW):
    """[N,C,H,W] -> columns [N, oH*oW, C*kH*kW] (K order = c, u, v like THNN unfolded_copy)."""
    N, C, H, W = x.shape
    oH, oW = conv_out_size(H, kH, dH, pH), conv_out_size(W, kW, dW, pW)
    xp = np.zeros((N, C, H + 2 * pH, W + 2 * pW), dtype=x.dtype)
    xp[:, :, pH:pH + H, pW:pW + W] = x
    cols = np.empty((N, oH, oW, C, kH, kW), dtype=x.dtype)
    for u in range(kH):
        for v in range(kW):
            cols[:, :, :, :, u, v] = xp[:, :, u:u + dH * oH:dH, v:v + dW * oW:dW].transpose(0, 2, 3, 1)
    return cols.reshape(N, oH * oW, C * kH * kW), oH, oW


def col2im(cols, C, H, W, kH, kW, dH, dW, pH, pW):
    """Adjoint of im2col: columns [N, oH*oW, C*kH*kW] -> [N,C,H,W] (accumulating overlaps)."""
    N = cols.shape[0]
    oH, oW = conv_out_size(H, kH, dH, pH), conv_out_size(W, kW, dW, pW)
    c6 = cols.reshape(N, oH, oW, C, kH, kW)
    xp = np.zeros((N, C, H + 2 * pH, W + 2 * pW), dtype=cols.dtype)
    for u in range(kH):
        for v in range(kW):
            xp[:, :, u:u + dH * oH:dH, v:v + dW * oW:dW] += c6[:, :, :, :, u, v].transpose(0, 3, 1, 2)
    return np.ascontiguousarray(xp[:, :, pH:pH + H, pW:pW + W])


# --------------------------------------------------------------------------
# SpatialConvolution (used at train.lua:89-104,183-196)
# --------------------------------------------------------------------------


def conv_forward(x, w, b, dH, dW, pH, pW):
    """y[n,o,i,j] = b[o] + sum_{c,u,v} w[o,c,u,v] x[n,c,i*d-p+u,j*d-p+v] (cross-correlation)."""
    Co, Ci, kH, kW = w.shape
    cols, oH, oW = im2col(x, kH, kW, dH, dW, pH, pW)
    y = cols @ w.reshape(Co, -1).T  # [N, oH*oW, Co]
    if b is not None:
        y = y + b
    return np.ascontiguousarray(y.transpose(0, 2, 1)).reshape(x.shape[0], Co, oH, oW)


def conv_grad_input(x_shape, gy, w, dH, dW, pH, pW):
    """gx = col2im(W^T gy) (SpatialConvolutionMM_updateGradInput)."""
    N, C, H, W = x_shape
    Co, Ci, kH, kW = w.shape
    g = gy.reshape(N, Co, -1).transpose(0, 2, 1)  # [N, oH*oW, Co]
    gcols = g @ w.reshape(Co, -1)  # [N, oH*oW, Ci*kH*kW]
    return col2im(gcols, C, H, W, kH, kW, dH, dW, pH, pW)


def conv_acc_grad(x, gy, gw, gb, dH, dW, pH, pW, scale=1.0):
    """gradWeight += scale * sum_n gy[n] im2col(x[n])^T ; gradBias += scale * sum gy (in place)."""
    Co, Ci, kH, kW = gw.shape
    N = x.shape[0]
    cols, oH, oW = im2col(x, kH, kW, dH, dW, pH, pW)
    g = gy.reshape(N, Co, -1)  # [N, Co, oH*oW]
    acc = np.einsum('nop,npk->ok', g, cols, optimize=True)
    gw += (scale * acc).reshape(gw.shape).astype(gw.dtype)
    if gb is not None:
        gb += (scale * g.sum(axis=(0, 2))).astype(gb.dtype)


# --------------------------------------------------------------------------
# SpatialFullConvolution (train.lua:134-146) -- weight [Cin,Cout,kH,kW]
# --------------------------------------------------------------------------


def fullconv_forward(x, w, b, dH, dW, pH, pW, adjH=0, adjW=0):
    """Transposed conv: scatter x[n,c,i,j] * w[c,o,u,v] to (i*d-p+u, j*d-p+v)."""
    N, Ci, H, W = x.shape
    Ci2, Co, kH, kW = w.shape
    assert Ci == Ci2
    oH, oW = fullconv_out_size(H, kH, dH, pH, adjH), fullconv_out_size(W, kW, dW, pW, adjW)
    # columns[n, p, (o,u,v)] = sum_c x[n,c,p] w[c,(o,u,v)]  then col2im into the output
    cols = x.reshape(N, Ci, H * W).transpose(0, 2, 1) @ w.reshape(Ci, -1)
    assert conv_out_size(oH, kH, dH, pH) == H and conv_out_size(oW, kW, dW, pW) == W
    y = col2im(cols, Co, oH, oW, kH, kW, dH, dW, pH, pW)
    if b is not None:
        y = y + b.reshape(1, Co, 1, 1)
    return y


def fullconv_grad_input(gy, w, dH, dW, pH, pW):
    """dgrad of the transposed conv = ordinary conv of gy with w (stride d, pad p)."""
    Ci, Co, kH, kW = w.shape
    cols, H, W = im2col(gy, kH, kW, dH, dW, pH, pW)  # [N, H*W, Co*kH*kW]
    gx = cols @ w.reshape(Ci, -1).T  # [N, H*W, Ci]
    return np.ascontiguousarray(gx.transpose(0, 2, 1)).reshape(gy.shape[0], Ci, H, W)


def fullconv_acc_grad(x, gy, gw, gb, dH, dW, pH, pW, scale=1.0):
    """gradWeight[c,o,u,v] += scale * sum_{n,i,j} x[n,c,i,j] gy[n,o,i*d-p+u,j*d-p+v]."""
    Ci, Co, kH, kW = gw.shape
    N = x.shape[0]
    cols, H, W = im2col(gy, kH, kW, dH, dW, pH, pW)  # [N, H*W, Co*kH*kW]
    xf = x.reshape(N, Ci, H * W)
    acc = np.einsum('ncp,npk->ck', xf, cols, optimize=True)
    gw += (scale * acc).reshape(gw.shape).astype(gw.dtype)
    if gb is not None:
        gb += (scale * gy.sum(axis=(0, 2, 3))).astype(gb.dtype)


# --------------------------------------------------------------------------
# SpatialBatchNormalization (train.lua:79)
# --------------------------------------------------------------------------


def bn_forward(x, gamma, beta, running_mean, running_var, train, momentum=0.1, eps=1e-5):
    """Returns (y, save_mean, save_invstd).  running_* are updated in place in train mode.

    THNN accumulates the sums in double for float tensors; we do the same.
    """
    N, C, H, W = x.shape
    n = N * H * W
    if train:
        xd = x.astype(np.float64)
        mean = xd.mean(axis=(0, 2, 3))
        S = ((xd - mean.reshape(1, C, 1, 1)) ** 2).sum(axis=(0, 2, 3))
        invstd = 1.0 / np.sqrt(S / n + eps)
        running_mean[:] = (momentum * mean + (1 - momentum) * running_mean).astype(running_mean.dtype)
        unbiased = S / (n - 1) if n > 1 else np.full_like(S, np.inf)
        running_var[:] = (momentum * unbiased + (1 - momentum) * running_var).astype(running_var.dtype)
    else:
        mean = running_mean.astype(np.float64)
        invstd = 1.0 / np.sqrt(running_var.astype(np.float64) + eps)
    mean = mean.astype(x.dtype)
    invstd = invstd.astype(x.dtype)
    g = gamma if gamma is not None else np.ones(C, x.dtype)
    b = beta if beta is not None else np.zeros(C, x.dtype)
    y = (x - mean.reshape(1, C, 1, 1)) * (invstd * g).reshape(1, C, 1, 1) + b.reshape(1, C, 1, 1)
    return y.astype(x.dtype), mean, invstd


def bn_backward(x, gy, gamma, save_mean, save_invstd, running_mean, running_var, train, eps=1e-5,
                ggamma=None, gbeta=None, scale=1.0, want_gx=True):
    """BatchNormalization_backward.  Accumulates into ggamma/gbeta when given; returns gx."""
    N, C, H, W = x.shape
    n = N * H * W
    g = (gamma if gamma is not None else np.ones(C, x.dtype)).reshape(1, C, 1, 1)
    if train:
        mean = save_mean.reshape(1, C, 1, 1)
        invstd = save_invstd.reshape(1, C, 1, 1)
    else:
        mean = running_mean.reshape(1, C, 1, 1).astype(x.dtype)
        invstd = (1.0 / np.sqrt(running_var.astype(np.float64) + eps)).astype(x.dtype).reshape(1, C, 1, 1)
    gyd = gy.astype(np.float64)
    s = gyd.sum(axis=(0, 2, 3))
    d = ((x.astype(np.float64) - mean) * gyd).sum(axis=(0, 2, 3))
    gx = None
    if want_gx:
        if train:
            sN = (s / n).astype(x.dtype).reshape(1, C, 1, 1)
            k = (d / n).astype(x.dtype).reshape(1, C, 1, 1) * invstd * invstd
            gx = (gy - sN - (x - mean) * k) * invstd * g
        else:
            gx = gy * invstd * g
        gx = gx.astype(x.dtype)
    if ggamma is not None:
        ggamma += (scale * d * invstd.reshape(C)).astype(ggamma.dtype)
    if gbeta is not None:
        gbeta += (scale * s).astype(gbeta.dtype)
    return gx


# --------------------------------------------------------------------------
# Activations
# --------------------------------------------------------------------------


def leaky_relu(x, negval=0.2):
    return np.where(x > 0, x, x * np.asarray(negval, x.dtype)).astype(x.dtype)


def leaky_relu_grad(x_or_y, gy, negval=0.2):
    """In-place LeakyReLU evaluates the mask on the overwritten tensor; sign-preserving so equivalent."""
    return np.where(x_or_y > 0, gy, gy * np.asarray(negval, gy.dtype)).astype(gy.dtype)


def relu(x):
    return np.where(x > 0, x, np.zeros((), x.dtype)).astype(x.dtype)


def relu_grad(x_or_y, gy):
    return np.where(x_or_y > 0, gy, np.zeros((), gy.dtype)).astype(gy.dtype)


def tanh(x):
    return np.tanh(x)


def tanh_grad(y, gy):
    return (gy * (1 - y * y)).astype(gy.dtype)


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(x.dtype)


def sigmoid_grad(y, gy):
    return (gy * y * (1 - y)).astype(gy.dtype)


# --------------------------------------------------------------------------
# Criteria
# --------------------------------------------------------------------------
BCE_EPS = 1e-12


def bce_forward(x, t):
    """nn.BCECriterion sizeAverage: -(1/n) sum t log(x+eps) + (1-t) log(1-x+eps); eps = 1e-12."""
    x = x.reshape(-1)
    t = t.reshape(-1)
    e = np.asarray(BCE_EPS, x.dtype)
    return float(-(t * np.log(x + e) + (1 - t) * np.log(1 - x + e)).sum(dtype=np.float64) / x.size)


def bce_backward(x, t):
    e = np.asarray(BCE_EPS, x.dtype)
    tt = t.reshape(x.shape)
    return (-(tt - x) / ((1 - x + e) * (x + e)) / x.size).astype(x.dtype)


def mse_forward(x, t):
    d = (x - t).astype(np.float64)
    return float((d * d).sum() / x.size)


def mse_backward(x, t):
    return ((x - t) * np.asarray(2.0 / x.size, x.dtype)).astype(x.dtype)


def abs_criterion_forward(x, t):
    return float(np.abs((x - t).astype(np.float64)).sum() / x.size)


def sgn_plus(z):
    """THNN Abs/AbsCriterion backward use (z >= 0 ? +1 : -1)."""
    return np.where(z >= 0, 1.0, -1.0).astype(z.dtype)


def abs_criterion_backward(x, t):
    return (sgn_plus(x - t) / x.size).astype(x.dtype)


def masked_mse_forward(x, t, mask, mW):
    """MaskedMSECriterion.lua:15-20,29-35: sum(((1-mW) M + mW) (x-t)^2) / n."""
    wM = mask.astype(np.float64) * (1 - mW) + mW
    d = (x - t).astype(np.float64)
    return float((wM * d * d).sum() / x.size)


def masked_mse_backward(x, t, mask, mW):
    """MaskedMSECriterion.lua:37-42: 2 wM (x-t) / n  (w.r.t. input only)."""
    wM = (mask.astype(x.dtype) * np.asarray(1 - mW, x.dtype) + np.asarray(mW, x.dtype))
    return (wM * (x - t) * np.asarray(2.0 / x.size, x.dtype)).astype(x.dtype)


def _gdl_terms(T):
    """Flat-index crops of gdl_criterion.lua:12-19 for one [B,C,H,W] tensor (H == W required)."""
    B, C, H, W = T.shape
    if H != W:
        raise ValueError("inconsistent tensor size (GDLCriterion needs square maps)")
    i1 = T[:, :, 0:H - 1, :].reshape(B, C, -1)  # drop bottom row   [H-1, W]
    j1 = T[:, :, 1:H, :].reshape(B, C, -1)      # drop top row      [H-1, W]
    i2 = T[:, :, :, 0:W - 1].reshape(B, C, -1)  # drop right column [H, W-1]
    j2 = T[:, :, :, 1:W].reshape(B, C, -1)      # drop left column  [H, W-1]
    return i1, j1, i2, j2


def gdl_forward(inp, target):
    """nn.GDLCriterion(1) (gdl_criterion.lua:38-45) with the flat-index pairing of SURVEY 9.8."""
    yi1, yj1, yi2, yj2 = _gdl_terms(target)
    hi1, hj1, hi2, hj2 = _gdl_terms(inp)
    t12 = np.abs(yi2 - yi1) - np.abs(hi2 - hi1)
    t34 = np.abs(yj2 - yj1) - np.abs(hj2 - hj1)
    n = t12.size
    return float(np.abs(t12.astype(np.float64)).sum() / n + np.abs(t34.astype(np.float64)).sum() / n)


def gdl_backward(inp, target):
    """gradInput w.r.t. ``inp`` (gdl_criterion.lua:47-52)."""
    B, C, H, W = inp.shape
    yi1, yj1, yi2, yj2 = _gdl_terms(target)
    hi1, hj1, hi2, hj2 = _gdl_terms(inp)
    n = yi1.size
    dt = inp.dtype
    g12 = sgn_plus(np.abs(yi2 - yi1) - np.abs(hi2 - hi1)) / np.asarray(n, dt)
    g34 = sgn_plus(np.abs(yj2 - yj1) - np.abs(hj2 - hj1)) / np.asarray(n, dt)
    d2 = -g12 * sgn_plus(hi2 - hi1)  # flat [H*(W-1)]
    d4 = -g34 * sgn_plus(hj2 - hj1)
    gx = np.zeros_like(inp)
    gx[:, :, :, 0:W - 1] += d2.reshape(B, C, H, W - 1)   # +d2 -> Yhat_i2 (pad right)
    gx[:, :, 0:H - 1, :] -= d2.reshape(B, C, H - 1, W)   # -d2 -> Yhat_i1 (pad bottom)
    gx[:, :, :, 1:W] += d4.reshape(B, C, H, W - 1)       # +d4 -> Yhat_j2 (pad left)
    gx[:, :, 1:H, :] -= d4.reshape(B, C, H - 1, W)       # -d4 -> Yhat_j1 (pad top)
    return gx.astype(dt)


# --------------------------------------------------------------------------
# Step glue (train.lua:377-400, train_vid_weighted.lua:485-528, inpaint_utils.lua:63-101)
# --------------------------------------------------------------------------


def overlap_weight_matrix(shape, overlapPred, wtl2, dtype):
    """wtl2Matrix of train.lua:390-392: 10*wtl2 on the border ring, wtl2 inside."""
    Wm = np.full(shape, 10.0 * wtl2, dtype=dtype)
    H, W = shape[2], shape[3]
    Wm[:, :, overlapPred:H - overlapPred, overlapPred:W - overlapPred] = wtl2
    return Wm


def blend_l2_overlap(df_dg, x, t, wtl2, overlapPred):
    """train.lua:377-400: df_dg <- (1-wtl2) df_dg + W .* df_l2 (0<wtl2<1) or df_dg + W .* df_l2."""
    df_l2 = mse_backward(x, t)
    dt = x.dtype
    if overlapPred == 0:
        Wm = np.asarray(wtl2, dt)
    else:
        Wm = overlap_weight_matrix(x.shape, overlapPred, wtl2, dt)
    if 0 < wtl2 < 1:
        return (df_dg * np.asarray(1 - wtl2, dt) + Wm * df_l2).astype(dt)
    return (df_dg + Wm * df_l2).astype(dt)


def blend_l2_masked(df_dg, x, t, mask01, wtl2, weight_nomask):
    """train_vid_weighted.lua:485-507 with overlapPred == 0.

    ``mask01`` is input_mask as float {0,1}.  Returns (new df_dg, weights) --
    the script overwrites input_mask with ``weights`` in place (:494).
    """
    dt = x.dtype
    df_l2 = mse_backward(x, t)
    weights = None
    if weight_nomask != 0:
        weights = (mask01 * np.asarray(1 - weight_nomask, dt) + np.asarray(weight_nomask, dt)).astype(dt)
        df_l2 = df_l2 * weights
    if 0 < wtl2 < 1:
        out = df_dg * np.asarray(1 - wtl2, dt) + np.asarray(wtl2, dt) * df_l2
    else:
        out = df_dg + np.asarray(wtl2, dt) * df_l2
    return out.astype(dt), weights


def mask_composite(dst, mask, src):
    """inpaint_utils.fillIn without scaling: dst[mask] = src[mask] (maskedSelect + maskedCopy)."""
    return np.where(mask != 0, src, dst).astype(dst.dtype)


def adam_step(x, g, state, lr, beta1, beta2=0.999, eps=1e-8):
    """optim.adam; ``state`` holds t, m, v; x updated in place."""
    if 't' not in state:
        state['t'] = 0
        state['m'] = np.zeros_like(x)
        state['v'] = np.zeros_like(x)
    state['t'] += 1
    t = state['t']
    dt = x.dtype
    m, v = state['m'], state['v']
    m *= np.asarray(beta1, dt)
    m += np.asarray(1 - beta1, dt) * g
    v *= np.asarray(beta2, dt)
    v += np.asarray(1 - beta2, dt) * g * g
    denom = np.sqrt(v) + np.asarray(eps, dt)
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    step = lr * np.sqrt(bc2) / bc1
    x -= (np.asarray(step, dt) * m / denom).astype(dt)
