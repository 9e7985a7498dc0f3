import sys; sys.path.insert(0, '.')
import numpy as np, asg_b200 as asg
asg.init(0)
f = lambda X: np.sum(np.sin(2*np.pi*X), axis=1)
# regular grid: stream kernel R=1 (small batch) and R=2 (large batch), masked, surplus, basis
G = asg.SparseGridInterpolant(4)
asg.train(G, f, 6, (4,5,6,3), vectorized=True)
rng = np.random.default_rng(0)
X = rng.random((330000, 4))
y = G.evaluate(X); y4 = G.evaluate(X, max_depth=4); s = G.surplus(X[:1000], y[:1000])
print('regular', len(G), float(np.abs(s).max()), G.info()['regular'])
# adaptive grid: hashed subspaces, generic groups
A = asg.SparseGridInterpolant(2)
u = lambda X: -1.0/(0.1+np.sum(X,axis=1))
asg.train(A, u, 4, (3,8), vectorized=True); asg.adapt(A, u, 12, tol=1e-4, verbose=False, vectorized=True)
ya = A.evaluate(rng.random((310000, 2))); B = asg.basis(rng.random((200,2)), A)
print('adaptive', len(A), B.nnz)
for p in ('trie','sweep'):
    A.set_path(p); A.evaluate(rng.random((5000,2)))
Bd = asg.basis(X[:100], G, dense=True); E = asg.dehierarchization_matrix(G)
print('ok', Bd.shape, E.nnz)
