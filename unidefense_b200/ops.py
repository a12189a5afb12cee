"""torch.autograd.Function wrappers over the C ABI (one per hot-path row of SURVEY.md §8a)."""
import torch

from . import _lib as L


class _ReconTail(torch.autograd.Function):
    """a1: model/unidefense.py:244-253 / :423-433 / :618-628."""

    @staticmethod
    def forward(ctx, dec, x, norm_ortho):
        dec = dec.contiguous()
        x = x.contiguous()
        L.require_cuda_f32(dec, x)
        N, C, h, w = dec.shape
        H, W = x.shape[-2:]
        if x.shape[0] != N or x.shape[1] != C:
            raise ValueError(f"recon_tail: dec {tuple(dec.shape)} and x {tuple(x.shape)} disagree")
        lib = L.lib()
        rec = torch.empty_like(x)
        spatial = torch.empty(N, device=x.device, dtype=torch.float32)
        freq = torch.empty(N, device=x.device, dtype=torch.float32)
        need_grad = ctx.needs_input_grad[0]
        signs = (torch.empty(lib.ud_recon_tail_signs_bytes(N, C, H, W), dtype=torch.uint8, device=x.device)
                 if need_grad else None)
        nws = lib.ud_recon_tail_workspace_bytes(N, C, h, w, H, W)
        ws = L.workspace(nws, x.device)
        L.check(lib.ud_recon_tail_fwd(L.ptr(dec), L.ptr(x), L.ptr(rec), L.ptr(spatial), L.ptr(freq), L.ptr(signs),
                                      L.ptr(ws), ws.numel(), N, C, h, w, H, W, int(norm_ortho), L.stream()),
                "recon_tail_fwd")
        if need_grad:
            ctx.save_for_backward(dec, x, signs)
        ctx.norm_ortho = int(norm_ortho)
        ctx.mark_non_differentiable(rec)
        return rec, spatial, freq

    @staticmethod
    def backward(ctx, g_rec, g_spatial, g_freq):
        dec, x, signs = ctx.saved_tensors
        N, C, h, w = dec.shape
        H, W = x.shape[-2:]
        lib = L.lib()
        gs = (g_spatial if g_spatial is not None else torch.zeros(N, device=x.device)).contiguous().float()
        gf = (g_freq if g_freq is not None else torch.zeros(N, device=x.device)).contiguous().float()
        g_dec = torch.empty_like(dec)
        nws = lib.ud_recon_tail_workspace_bytes(N, C, h, w, H, W)
        ws = L.workspace(nws, x.device)
        L.check(lib.ud_recon_tail_bwd(L.ptr(dec), L.ptr(x), L.ptr(signs), L.ptr(gs), L.ptr(gf), L.ptr(g_dec),
                                      L.ptr(ws), ws.numel(), N, C, h, w, H, W, ctx.norm_ortho, L.stream()),
                "recon_tail_bwd")
        return g_dec, None, None


def recon_tail(dec, x, norm="ortho"):
    """-> (rec [N,C,H,W], spatial [N], freq [N]); `rec` carries no grad (it is only returned for
    visualisation/eval by the reference, engine/forgery_engine.py:343-347)."""
    return _ReconTail.apply(dec, x, norm == "ortho")
