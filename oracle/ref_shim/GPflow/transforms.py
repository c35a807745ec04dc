"""Transforms are carried as tags only: the stand-in stores CONSTRAINED values and differentiates with
respect to them (the chain rule to the free state is host-side bookkeeping, not on the hot path)."""


class Identity:
    pass


class positive(Identity):  # GPflow 0.x: softplus(x) + 1e-6
    pass


class LowerTriangular(Identity):
    def __init__(self, N, num_matrices=1, squeeze=False):
        self.N, self.num_matrices = N, num_matrices
