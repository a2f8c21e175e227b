import numpy as np, scipy.optimize as spo, time
def fit_plane_ref(image):
    lxx, lyy = np.meshgrid(np.arange(image.shape[0]), np.arange(image.shape[1]), indexing='ij')
    f = lambda x, im, xx, yy: (im - (x[0]*xx + x[1]*yy + x[2])).flatten()
    return spo.least_squares(f, np.zeros(3), loss='huber', args=(image, lxx, lyy))
def fit_plane_irls(image, tol=1e-11, maxit=500):
    n, m = image.shape
    x = np.arange(n)[:, None] - (n-1)/2; y = np.arange(m)[None, :] - (m-1)/2
    th = np.zeros(3); w = np.ones_like(image)
    for it in range(maxit):
        A = np.array([[ (w*x*x).sum(), (w*x*y).sum(), (w*x).sum()], [(w*x*y).sum(), (w*y*y).sum(), (w*y).sum()], [(w*x).sum(), (w*y).sum(), w.sum()]])
        b = np.array([(w*x*image).sum(), (w*y*image).sum(), (w*image).sum()])
        new = np.linalg.solve(A, b)
        d = max(abs(new[0]-th[0])*n/2, abs(new[1]-th[1])*m/2, abs(new[2]-th[2]))
        th = new
        r = image - (th[0]*x + th[1]*y + th[2])
        w = 1/np.maximum(1.0, np.abs(r))
        if d < tol: break
    return np.array([th[0], th[1], th[2] - th[0]*(n-1)/2 - th[1]*(m-1)/2]), it+1
rng = np.random.default_rng(0)
for shape, slope, noise, outl in [((64,64),0.05,0.3,0.0), ((128,96),0.4,0.5,0.05), ((256,256),0.02,2.0,0.1), ((200,150), 1.5, 0.1, 0.2)]:
    n, m = shape
    xx, yy = np.meshgrid(np.arange(n), np.arange(m), indexing='ij')
    img = slope*xx - 0.7*slope*yy + 3 + noise*rng.normal(size=shape)
    mask = rng.uniform(size=shape) < outl
    img[mask] += rng.normal(size=mask.sum())*20
    t=time.time(); r = fit_plane_ref(img); t1=time.time()-t
    th, its = fit_plane_irls(img)
    print(shape, r.x, th, its, np.abs(r.x-th), r.nfev, f"{t1:.2f}s", r.status)
