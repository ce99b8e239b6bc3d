"""Shared helpers for the parity tests."""
import torch


def report_mismatch(name, got, ref, atol, rtol):
    """Return (ok, message) with enough structure to diagnose layout / swizzle / descriptor bugs remotely."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    if got.shape != ref.shape:
        return False, "%s: shape %s vs %s" % (name, tuple(got.shape), tuple(ref.shape))
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    nbad = int(bad.sum())
    finite = bool(torch.isfinite(got).all())
    msg = "%s: max_err=%.4g max_ref=%.4g bad=%d/%d finite=%s" % (name, float(err.max()), float(ref.abs().max()), nbad,
                                                                 err.numel(), finite)
    if nbad:
        flat = bad.reshape(-1, bad.shape[-1]) if bad.dim() > 1 else bad.reshape(1, -1)
        rows = flat.any(1).nonzero().flatten()
        cols = flat.any(0).nonzero().flatten()
        msg += " | bad rows %d/%d (first %s) bad cols %d/%d (first %s)" % (
            rows.numel(), flat.shape[0], rows[:12].tolist(), cols.numel(), flat.shape[1], cols[:12].tolist())
        i = int(err.reshape(-1).argmax())
        msg += " | worst got=%.5g ref=%.5g" % (float(got.reshape(-1)[i]), float(ref.reshape(-1)[i]))
    return nbad == 0 and finite, msg


def assert_close(name, got, ref, atol, rtol):
    ok, msg = report_mismatch(name, got, ref, atol, rtol)
    print(msg)
    assert ok, msg
