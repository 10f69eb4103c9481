"""shared helper of the perf probes: synthetic cylinder-lattice potential built with torch ops"""
import torch


def make_V(ctx, pitch=8.0, r=0.5):
    X = torch.from_numpy(ctx.X).cuda(); Y = torch.from_numpy(ctx.Y).cuda()
    fx = torch.remainder(X + pitch / 2, pitch) - pitch / 2
    fy = torch.remainder(Y + pitch / 2, pitch) - pitch / 2
    V = torch.zeros(ctx.Ny, ctx.Nx, dtype=torch.float64, device="cuda")
    V[(fy[:, None] ** 2 + fx[None, :] ** 2).sqrt() < r] = -100.0
    V[0, :] = -100; V[-1, :] = -100; V[:, 0] = -100; V[:, -1] = -100
    dfx = (torch.remainder(X + 32, 64.0) - 32).abs() < 1.0
    dfy = (torch.remainder(Y + 32, 64.0) - 32).abs() < 1.0
    V[dfy[:, None] & dfx[None, :]] = 1.0
    return V

