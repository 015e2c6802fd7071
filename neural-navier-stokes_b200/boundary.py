"""Boundary-condition objects: the reference's ``src/boundary.py`` API (same class names,
constructor signature, attributes ``value/boundary/dx/dy/type`` and asserts, boundary.py:14-23).

In the reference ``apply(A)`` is where the numerics happen.  Here the objects are descriptors:
the solver classes flatten their ``u_bc / v_bc / p_bc`` lists, IN LIST ORDER, into the ``nns_bc``
table of the C ABI and the edges are written inside the CUDA kernels.  ``apply`` is kept for
callers that use it directly on host arrays (e.g. to prepare initial conditions); it is a
host-side convenience on a numpy array, never part of a solver step.
"""

_SIDES = ('left', 'right', 'bottom', 'top')

# side -> (edge index, neighbour index, axis-0 edge?, sign of the one-sided difference)
# boundary.py:39-46 (Dirichlet) and :73-84 (Neumann)
_EDGE = {
    'left': (0, 1, True, -1.0),
    'right': (-1, -2, True, +1.0),
    'bottom': (0, 1, False, -1.0),
    'top': (-1, -2, False, +1.0),
}


class BaseBoundaryCondition(object):
    """value: Dirichlet value or normal derivative; boundary: 'left' (A[0,:]), 'right' (A[-1,:]),
    'bottom' (A[:,0]), 'top' (A[:,-1]); dx, dy: grid spacings (must be float)."""

    type = None

    def __init__(self, value, boundary, dx, dy):
        assert isinstance(boundary, str)
        assert isinstance(dx, float)
        assert isinstance(dy, float)
        assert boundary in _SIDES
        self.value, self.boundary = value, boundary
        self.dx, self.dy = dx, dy

    def apply(self, A):
        raise NotImplementedError

    # -- C-ABI descriptor ----------------------------------------------------------------
    def abi_codes(self):
        """(side, type) integer codes of include/nns_b200.h."""
        return _SIDES.index(self.boundary), 0 if self.type == 'dirichlet' else 1

    def __repr__(self):
        return "%s(%r, %r, %r, %r)" % (type(self).__name__, self.value, self.boundary, self.dx, self.dy)


class DirichletBoundaryCondition(BaseBoundaryCondition):
    def __init__(self, value, boundary, dx, dy):
        super().__init__(value, boundary, dx, dy)
        self.type = 'dirichlet'

    def apply(self, A):
        edge, _, axis0, _ = _EDGE[self.boundary]
        if axis0:
            A[edge, :] = self.value
        else:
            A[:, edge] = self.value
        return A


class NeumannBoundaryCondition(BaseBoundaryCondition):
    def __init__(self, value, boundary, dx, dy):
        super().__init__(value, boundary, dx, dy)
        self.type = 'neumann'

    def apply(self, A):
        edge, inner, axis0, sign = _EDGE[self.boundary]
        if axis0:
            A[edge, :] = A[inner, :] + sign * self.dx * self.value
        else:
            A[:, edge] = A[:, inner] + sign * self.dy * self.value
        return A
