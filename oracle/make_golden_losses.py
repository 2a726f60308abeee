"""TEST INFRASTRUCTURE ONLY — loss fixtures (filled in with the loss kernels)."""


def main(ref):
    return {}
