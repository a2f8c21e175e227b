import numpy as np, scipy.optimize as spo
def fit(image, tol=1e-11, maxit=500, c=1.0):
    n, m = image.shape
    x = ((np.arange(n) - (n-1)/2)/n)[:, None]*np.ones((1,m)); y = ((np.arange(m) - (m-1)/2)/m)[None, :]*np.ones((n,1))
    th = np.zeros(3); acc=None; f_acc=None; irls=None; was_newton=False; it=0
    while True:
        first = it == 0
        res = image - (th[0]*x + th[1]*y + th[2]); ares = np.abs(res)
        inl = np.ones_like(res, bool) if first else ares <= c
        w = np.where(inl, 1.0, c/np.maximum(ares,1e-300))
        F = np.where(inl, 0.5*res*res, c*ares-0.5*c*c).sum()
        it += 1
        if not first and was_newton and not (F <= f_acc):
            th = irls.copy(); was_newton=False
            if it>=maxit: return acc, it
            continue
        acc = th.copy(); f_acc = F
        A = np.array([[(w*x*x).sum(),(w*x*y).sum(),(w*x).sum()],[(w*x*y).sum(),(w*y*y).sum(),(w*y).sum()],[(w*x).sum(),(w*y).sum(),w.sum()]])
        b = np.array([(w*x*image).sum(),(w*y*image).sum(),(w*image).sum()])
        irls = np.linalg.solve(A,b); nxt = irls.copy(); newton=False
        if not first and inl.sum()>=3:
            H = np.array([[(inl*x*x).sum(),(inl*x*y).sum(),(inl*x).sum()],[(inl*x*y).sum(),(inl*y*y).sum(),(inl*y).sum()],[(inl*x).sum(),(inl*y).sum(),inl.sum()]])
            psi = w*res; g = np.array([(psi*x).sum(),(psi*y).sum(),psi.sum()])
            try:
                nxt = th + np.linalg.solve(H,g); newton=True
            except np.linalg.LinAlgError: pass
        d = 0.5*abs(nxt[0]-th[0])+0.5*abs(nxt[1]-th[1])+abs(nxt[2]-th[2])
        was_newton=newton; th = nxt
        if (not first and d<tol) or it>=maxit:
            xc,yc=(n-1)/2,(m-1)/2; a0=th[0]/n; a1=th[1]/m
            return np.array([a0,a1,th[2]-a0*xc-a1*yc]), it
def ref(image):
    xx, yy = np.meshgrid(np.arange(image.shape[0]), np.arange(image.shape[1]), indexing='ij')
    return spo.least_squares(lambda p: (image-(p[0]*xx+p[1]*yy+p[2])).ravel(), np.zeros(3), loss='huber').x
rng=np.random.default_rng(0)
cases=[]
for shape,slope,noise,outl in [((64,64),0.05,0.3,0.0),((128,96),0.4,0.5,0.05),((257,130),0.02,2.0,0.1),((200,150),1.5,0.1,0.2)]:
    n,m=shape; xx,yy=np.meshgrid(np.arange(n),np.arange(m),indexing='ij')
    img=slope*xx-0.7*slope*yy+3+noise*rng.normal(size=shape); mask=rng.uniform(size=shape)<outl; img[mask]+=rng.normal(size=mask.sum())*20
    cases.append(img)
# displacement-like smooth field with large residuals
n=m=256; xx,yy=np.meshgrid(np.arange(n),np.arange(m),indexing='ij')
cases.append(8*np.exp(-((xx-100)**2+(yy-140)**2)/3000.0)+0.01*xx)
cases.append(30*np.sin(xx/40.0)*np.cos(yy/55.0)+0.2*yy)
g=dict(np.load('tests/golden/iterate_96x80.npz')); cases.append(g['in_plane'])
for img in cases:
    th,it=fit(img); r=ref(img); print(img.shape, it, np.abs(th-r).max(), th)
