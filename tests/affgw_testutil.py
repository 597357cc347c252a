"""Small comparison helpers shared by the test modules (kept out of conftest so that they can be imported by name)."""


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def cosine(a, b):
    a, b = a.detach().double().reshape(-1).cpu(), b.detach().double().reshape(-1).cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))
