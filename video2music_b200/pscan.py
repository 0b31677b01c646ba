"""Drop-in for model/pscan.py:228 (`pscan = PScan.apply`): H[t] = A[t] * H[t-1] + X[t] along dim 1 of
(B, L, D, N) fp32 tensors, with autograd.  Inputs are not modified (the reference clones, pscan.py:170-176)."""
import torch

from . import ops


class PScan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A_in, X_in):
        if A_in.dim() != 4 or A_in.shape != X_in.shape:
            raise ValueError("pscan expects A and X of identical shape (B, L, D, N), got %s and %s"
                             % (tuple(A_in.shape), tuple(X_in.shape)))
        A = A_in.detach().float().contiguous()
        X = X_in.detach().float().contiguous()
        H = ops.pscan_fwd(A, X)
        ctx.save_for_backward(A, H)
        return H

    @staticmethod
    def backward(ctx, grad_output):
        A, H = ctx.saved_tensors
        gA, gX = ops.pscan_bwd(A, H, grad_output.float().contiguous())
        return gA, gX


pscan = PScan.apply
