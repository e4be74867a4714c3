"""Stand-in for torchdiffeq.odeint — fixed-grid solvers only (reference call: e2_tts.py:1421,
method='midpoint'). Grid = the `t` tensor itself; atol/rtol are ignored by fixed-grid methods.
Test infrastructure only."""
import torch


def odeint(fn, y0, t, *, method='midpoint', atol=None, rtol=None, **kw):
    assert method in ('midpoint', 'euler'), method
    ys = [y0]
    y = y0
    for i in range(t.shape[0] - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        f0 = fn(t0, y)
        if method == 'euler':
            y = y + dt * f0
        else:
            half = 0.5 * dt
            ymid = y + f0 * half
            y = y + dt * fn(t0 + half, ymid)
        ys.append(y)
    return torch.stack(ys)
