"""Algorithm-level checks of device code that has no CPU twin in the product: numpy MIRRORS of kernel algorithms (same steps, same
tolerances) against brute force.  They do not execute CUDA; they establish that the algorithm the kernel implements is right over
configurations the GPU parity tests (random hand poses) do not sample.

capsule-ellipsoid collider (myo_device.cuh: collide_ellipsoid, CT_CAP_ELL): distance to the capsule's axis LINE as a 1-D Newton on the
circle normal to the axis, then -- when the line's closest point lies beyond a cap -- the end point's distance (Newton on the sphere,
ell_sd with a point).  Brute force: point-ellipsoid distance by its Lagrange root, minimised over the segment (scipy)."""
import numpy as np
from scipy.optimize import brentq, minimize_scalar
def pt_ell(y,a):
    f=lambda t: np.sum((a*y/(a*a+t))**2)-1.0
    if f(0.0)<=0: return None
    hi=1.0
    while f(hi)>0: hi*=4
    t=brentq(f,0.0,hi,xtol=1e-18,rtol=1e-15,maxiter=500); x=a*a*y/(a*a+t); return np.linalg.norm(y-x)
def brute(c1,a1,h,c2,R2,s2):
    f=lambda t: pt_ell(R2.T@(c1+a1*t-c2), s2)
    vals=[f(t) for t in np.linspace(-h,h,41)]
    if any(v is None for v in vals): return None
    res=minimize_scalar(f,bounds=(-h,h),method="bounded",options=dict(xatol=1e-13))
    return min(res.fun,f(-h),f(h))
def ell_sd_point(dl,R2,s2,d):
    # Newton on sphere for point vs ellipsoid (mirror of kernel's ell_sd with s1 = None)
    f=0
    for it in range(60):
        a=R2.T@d; u=s2*s2*a; n2=np.sqrt(a@u); p2=R2@u/n2
        g=dl-p2; f=d@dl-n2
        e=np.zeros(3); e[np.argmin(np.abs(d))]=1
        t1=np.cross(d,e); t1/=np.linalg.norm(t1); t2=np.cross(d,t1)
        g1,g2=t1@g,t2@g; scale=np.linalg.norm(dl)+n2
        if g1*g1+g2*g2<1e-24*scale*scale: break
        b1,b2=R2.T@t1,R2.T@t2; v=s2*s2
        a11=np.sum(v*b1*b1); a12=np.sum(v*b1*b2); a22=np.sum(v*b2*b2)
        q1,q2=t1@p2,t2@p2
        H11=-f-(a11-q1*q1)/n2; H12=-(a12-q1*q2)/n2; H22=-f-(a22-q2*q2)/n2
        det=H11*H22-H12*H12
        if H11<0 and det>0: dx=-(H22*g1-H12*g2)/det; dy=-(-H12*g1+H11*g2)/det
        else:
            L=abs(H11)+abs(H22)+abs(H12)+1e-12; dx=g1/L; dy=g2/L
        nn=np.hypot(dx,dy)
        if nn>0.5: dx*=0.5/nn; dy*=0.5/nn
        last=nn<1e-12
        for bt in range(12):
            dn=d+dx*t1+dy*t2; dn/=np.linalg.norm(dn)
            aa=R2.T@dn; fn=dn@dl-np.sqrt(np.sum(s2*s2*aa*aa))
            if fn>=f-1e-14*scale or bt==11: d=dn; break
            dx*=0.5; dy*=0.5
        if last: break
    return f,d
def kernel_alg(c1,a1,h,c2,R2,s2):
    dv=c2-c1
    e=np.zeros(3); e[np.argmin(np.abs(a1))]=1
    e1=np.cross(a1,e); e1/=np.linalg.norm(e1); e2=np.cross(a1,e1)
    b1,b2=R2.T@e1,R2.T@e2; v=s2*s2
    A11=np.sum(v*b1*b1); A12=np.sum(v*b1*b2); A22=np.sum(v*b2*b2)
    cc1,cc2=e1@dv,e2@dv; cn=np.hypot(cc1,cc2); u1,u2=(cc1/cn,cc2/cn) if cn>1e-15 else (1.0,0.0)
    scale=cn+np.sqrt(max(A11,A22)); F=0
    for it in range(40):
        Au1=A11*u1+A12*u2; Au2=A12*u1+A22*u2; n2=u1*Au1+u2*Au2; n=np.sqrt(n2)
        F=cc1*u1+cc2*u2-n
        p1,p2=-u2,u1
        uAp=p1*Au1+p2*Au2; pAp=A11*p1*p1+2*A12*p1*p2+A22*p2*p2
        g=cc1*p1+cc2*p2-uAp/n; H=-(cc1*u1+cc2*u2)-((pAp-n2)/n-uAp*uAp/n**3)
        if abs(g)<1e-12*scale: break
        dx=-g/H if H<0 else g/(abs(H)+1e-12)
        dx=max(-0.5,min(0.5,dx)); last=abs(dx)<1e-12
        for bt in range(12):
            w1,w2=u1+dx*p1,u2+dx*p2; q=1/np.hypot(w1,w2); w1*=q; w2*=q
            fn=cc1*w1+cc2*w2-np.sqrt(A11*w1*w1+2*A12*w1*w2+A22*w2*w2)
            if fn>=F-1e-14*scale or bt==11: u1,u2=w1,w2; break
            dx*=0.5
        if last: break
    d=u1*e1+u2*e2
    a=R2.T@d; u=v*a; nn=np.sqrt(a@u); p2v=R2@u/nn
    t=(dv-p2v)@a1
    if -h<=t<=h: return F
    t=h if t>h else -h
    sd,_=ell_sd_point(dv-a1*t,R2,s2,d)
    return sd
def randrot():
    q=rng.normal(size=4); q/=np.linalg.norm(q); w,x,y,z=q
    return np.array([[1-2*(y*y+z*z),2*(x*y-z*w),2*(x*z+y*w)],[2*(x*y+z*w),1-2*(x*x+z*z),2*(y*z-x*w)],[2*(x*z-y*w),2*(y*z+x*w),1-2*(x*x+y*y)]])


def test_capsule_ellipsoid_algorithm_vs_brute_force():
    global rng
    rng = np.random.default_rng(0)
    n = ncap = 0; worst = 0.0
    for trial in range(500):
        s2 = rng.uniform(0.3, 2.0, 3); R2 = randrot(); c2 = np.zeros(3)
        a1 = rng.normal(size=3); a1 /= np.linalg.norm(a1); h = rng.uniform(0.05, 3.0)
        c1 = rng.normal(size=3) * rng.uniform(0.5, 4.0)
        ref = brute(c1, a1, h, c2, R2, s2)
        if ref is None or ref < 1e-3:
            continue                                   # axis touches / enters the ellipsoid: outside the collider's specified regime
        got = kernel_alg(c1, a1, h, c2, R2, s2); n += 1
        worst = max(worst, abs(got - ref))
    assert n > 300 and worst < 1e-9, (n, worst)


def ell_sd_pair(dl, R1, s1, R2, s2, d):
    """Mirror of ell_sd (myo_device.cuh) for two ellipsoids: Newton on the unit sphere for max_d d.dl - h1(d) - h2(d)."""
    f = 0
    for it in range(60):
        a = R1.T @ d; u = s1 * s1 * a; n1 = np.sqrt(a @ u); p1 = R1 @ u / n1
        a = R2.T @ d; u = s2 * s2 * a; n2 = np.sqrt(a @ u); p2 = R2 @ u / n2
        g = dl - p1 - p2; f = d @ dl - n1 - n2
        e = np.zeros(3); e[np.argmin(np.abs(d))] = 1
        t1 = np.cross(d, e); t1 /= np.linalg.norm(t1); t2 = np.cross(d, t1)
        g1, g2 = t1 @ g, t2 @ g; scale = np.linalg.norm(dl) + n1 + n2
        if g1 * g1 + g2 * g2 < 1e-24 * scale * scale:
            break
        H11, H12, H22 = -f, 0.0, -f
        for R, s, p, nrm in ((R1, s1, p1, n1), (R2, s2, p2, n2)):
            b1, b2 = R.T @ t1, R.T @ t2; v = s * s
            a11, a12, a22 = np.sum(v * b1 * b1), np.sum(v * b1 * b2), np.sum(v * b2 * b2)
            q1, q2 = t1 @ p, t2 @ p
            H11 -= (a11 - q1 * q1) / nrm; H12 -= (a12 - q1 * q2) / nrm; H22 -= (a22 - q2 * q2) / nrm
        det = H11 * H22 - H12 * H12
        if H11 < 0 and det > 0:
            dx = -(H22 * g1 - H12 * g2) / det; dy = -(-H12 * g1 + H11 * g2) / det
        else:
            L = abs(H11) + abs(H22) + abs(H12) + 1e-12; dx = g1 / L; dy = g2 / L
        nn = np.hypot(dx, dy)
        if nn > 0.5:
            dx *= 0.5 / nn; dy *= 0.5 / nn
        last = nn < 1e-12
        for bt in range(12):
            dn = d + dx * t1 + dy * t2; dn /= np.linalg.norm(dn)
            fn = dn @ dl - np.sqrt(np.sum((s1 * (R1.T @ dn)) ** 2)) - np.sqrt(np.sum((s2 * (R2.T @ dn)) ** 2))
            if fn >= f - 1e-14 * scale or bt == 11:
                d = dn; break
            dx *= 0.5; dy *= 0.5
        if last:
            break
    return f


def test_ellipsoid_ellipsoid_algorithm_vs_brute_force():
    """ell_sd for two ellipsoids against a direct minimisation of the point-ellipsoid distance over the other ellipsoid's surface."""
    from scipy.optimize import minimize
    global rng
    rng = np.random.default_rng(1)
    n = 0; worst = 0.0
    for trial in range(40):
        s1, s2 = rng.uniform(0.3, 1.5, 3), rng.uniform(0.3, 1.5, 3); R1, R2 = randrot(), randrot()
        dl = rng.normal(size=3); dl *= rng.uniform(2.5, 5.0) / np.linalg.norm(dl)          # centre offset c2 - c1 (separated)
        surf = lambda ang: R1 @ (s1 * np.array([np.cos(ang[0]) * np.cos(ang[1]), np.sin(ang[0]) * np.cos(ang[1]), np.sin(ang[1])]))
        obj = lambda ang: pt_ell(R2.T @ (surf(ang) - dl), s2)
        _, a0, b0 = min((obj((a, b)), a, b) for a in np.linspace(0, 2 * np.pi, 40) for b in np.linspace(-np.pi / 2, np.pi / 2, 21))
        best = minimize(obj, (a0, b0), method="Nelder-Mead", options=dict(xatol=1e-12, fatol=1e-15, maxiter=4000)).fun     # grid start: no local traps
        got = ell_sd_pair(dl, R1, s1, R2, s2, dl / np.linalg.norm(dl)); n += 1
        worst = max(worst, abs(got - best))
    assert n == 40 and worst < 1e-9, worst
