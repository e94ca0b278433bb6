"""Problem construction helpers (reference: qradient/optimization_problems.py:4-37).

Host-side input construction only: a MaxCut instance becomes the observable dictionary
{'zz': upper-triangular array with 1.0 on every edge} that `Qaoa` consumes.
"""
import numpy as np


class MaxCut:
    def __init__(self, vertex_num, **kwargs):
        self.vertex_num = vertex_num
        if 'edge_set' in kwargs:
            self.edge_set = kwargs['edge_set']
        elif 'edge_num' in kwargs:
            self._random(edge_num=kwargs['edge_num'])
        elif 'edge_probability' in kwargs:
            self._random(edge_probability=kwargs['edge_probability'])
        else:
            raise ValueError('Specify one of the three edge_set, edge_num, or edge_probability')

    def _random(self, edge_num=None, edge_probability=None):
        pairs = np.array([[i, j] for i in range(self.vertex_num) for j in range(i + 1, self.vertex_num)])
        if edge_num is not None:
            # same draw as optimization_problems.py:22 (global numpy stream)
            chosen = np.random.choice(range(len(pairs)), size=edge_num, replace=False)
            self.edge_set = pairs[chosen]
        else:
            self.edge_set = np.array([e for e in pairs if np.random.rand() < edge_probability])

    def to_observable(self):
        zz = np.full([self.vertex_num, self.vertex_num], None)
        for a, b in self.edge_set:
            zz[a, b] = 1.
        return {'zz': zz}

    @staticmethod
    def random_regular(vertex_num, degree, seed):
        """Seeded pairing-model d-regular graph (extension; the reference has no regular-graph
        generator).  Returns a sorted edge list."""
        rng = np.random.default_rng(seed)
        if (vertex_num * degree) % 2:
            raise ValueError('vertex_num * degree must be even')
        while True:
            stubs = np.repeat(np.arange(vertex_num), degree)
            rng.shuffle(stubs)
            edges = set()
            ok = True
            for a, b in stubs.reshape(-1, 2):
                a, b = int(min(a, b)), int(max(a, b))
                if a == b or (a, b) in edges:
                    ok = False
                    break
                edges.add((a, b))
            if ok:
                return sorted(edges)
