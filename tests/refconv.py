"""CPU reference helpers for the kernel-level GPU tests (torch fp32/fp64 on the host)."""
import torch
import torch.nn.functional as F


def pack_fwd(w):
    """torch weight [c_out, cin_g, k] -> forward pack [k, c_out, cin_g]."""
    return w.permute(2, 0, 1).contiguous()


def pack_dgrad(w, groups):
    """torch weight [c_out, cin_g, k] -> data-gradient pack [k, c_in, cout_g]."""
    c_out, cin_g, k = w.shape
    cout_g = c_out // groups
    return w.view(groups, cout_g, cin_g, k).permute(3, 0, 2, 1).reshape(k, groups * cin_g, cout_g).contiguous()


def expand_groups(w, groups, pg):
    """torch weight [c_out, c_in/groups, k] of a `groups`-grouped conv -> the block-diagonal weight
    [c_out, c_in/pg, k] of the same conv written with pg | groups groups (zeros between the real groups)."""
    if pg == groups:
        return w
    c_out, cin_g, k = w.shape
    f = groups // pg
    cout_g = c_out // groups
    out = torch.zeros(c_out, cin_g * f, k, dtype=w.dtype)
    for g in range(groups):
        lo = (g % f) * cin_g
        out[g * cout_g:(g + 1) * cout_g, lo:lo + cin_g] = w[g * cout_g:(g + 1) * cout_g]
    return out


def unfold_ref(x_cl, *, phases, k, dilation, stride, pad, t_out):
    """im2col rows [B, t_out*phases, roundup8(k*C)] of a channels-last period view (host reference of stg_unfold)."""
    B, TP, Cc = x_cl.shape
    T = TP // phases
    x = x_cl.view(B, T, phases, Cc)
    kp = (k * Cc + 7) // 8 * 8
    out = torch.zeros(B, t_out, phases, kp, dtype=x_cl.dtype)
    for j in range(k):
        for t in range(t_out):
            ts = t * stride + j * dilation - pad
            if 0 <= ts < T:
                out[:, t, :, j * Cc:(j + 1) * Cc] = x[:, ts]
    return out.view(B, t_out * phases, kp)


def to_virtual(x_cl, phases):
    """[B, T*p, C] channels-last period view -> [B*p, C, T] torch conv1d layout."""
    B, TP, Cc = x_cl.shape
    T = TP // phases
    return x_cl.view(B, T, phases, Cc).permute(0, 2, 3, 1).reshape(B * phases, Cc, T)


def from_virtual(y, B, phases):
    """[B*p, C, T] -> [B, T*p, C]."""
    BP, Cc, T = y.shape
    return y.view(B, phases, Cc, T).permute(0, 3, 1, 2).reshape(B, T * phases, Cc).contiguous()


def conv_ref(x_cl, w, bias, *, phases=1, stride=1, dilation=1, pad=0, groups=1):
    """Channels-last reference conv (fp64 accumulate): x_cl [B, T*p, C_in] -> [B, T_out*p, C_out]."""
    B = x_cl.shape[0]
    xv = to_virtual(x_cl.double(), phases)
    y = F.conv1d(xv, w.double(), None if bias is None else bias.double(), stride=stride, padding=pad,
                 dilation=dilation, groups=groups)
    return from_virtual(y, B, phases)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
