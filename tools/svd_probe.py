import numpy as np
def props(u, s, v, refangle=0., refscale=1., diff=False):
    signs = np.sign(u[..., None, [0, 1], [0, 1]])
    v = signs*v
    u = np.swapaxes(signs*u, -1, -2)
    u_p = np.swapaxes(u @ v, -1, -2)
    angle = np.rad2deg(np.arctan2(u_p[..., 1, 0], u_p[..., 0, 0]))
    aniangle = np.rad2deg(np.arctan2(u[..., 1, 0], u[..., 0, 0]))
    if diff:
        aniangle += 90; alpha = s[..., 0]
    else:
        alpha = s[..., 1]
    kappa = s[..., 0] / s[..., 1]
    aniangle = aniangle % 180
    return np.array([angle + refangle, aniangle, alpha * refscale, kappa])
rng=np.random.default_rng(0)
A = np.eye(2)+0.2*rng.normal(size=(20000,2,2))
u,s,v = np.linalg.svd(A)
print("det u:", np.unique(np.sign(np.linalg.det(u)), return_counts=True))
print("det v:", np.unique(np.sign(np.linalg.det(v)), return_counts=True))
p0 = props(u,s,v)
# alternative: flip second pair
D = np.array([1.,-1.])
u2 = u*D[None,None,:]; v2 = v*D[None,:,None]
assert np.allclose(u2*s[:,None,:]@v2, A)
p1 = props(u2,s,v2)
print("max diff when flipping one pair:", np.abs(p0-p1).max(axis=1))
print(p0[:, :3].T); print(p1[:, :3].T)
