"""No-op stand-in for jax_tqdm (TEST INFRASTRUCTURE ONLY): ``@scan_tqdm(n)`` leaves the scan body untouched."""


def scan_tqdm(n, *a, **k):
    return lambda f: f


loop_tqdm = scan_tqdm
