"""NumPy restatement of cv::FarnebackOpticalFlow::calc as the reference calls it with `-f`
(flow.cpp:22-26).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

OpenCV is an un-vendored dependency of the reference; this restates its published algorithm
(modules/video/src/optflowgf.cpp: FarnebackPrepareGaussian / PolyExp / UpdateMatrices /
UpdateFlow_Blur, plus cv::GaussianBlur and cv::resize INTER_LINEAR on float images) and is pinned
against the cv2 4.13 binary in tests/test_oracle_cv.py (agreement ~2e-6 px).  The CUDA kernels in
csrc/farneback.cu are written from these formulas; the parity oracle itself is cv2
(oracle/flow.py::calculate_flow(use_farneback=True))."""
import numpy as np, cv2
f32=np.float32
def resize_linear(src, W, H):
    h,w=src.shape[:2]
    sx=w/W; sy=h/H
    def coeffs(n_dst,n_src,scale):
        d=np.arange(n_dst)
        f=((d+0.5)*scale-0.5).astype(f32)   # computed in double then float? cv: fx=(float)((dx+0.5)*scale_x-0.5)
        s=np.floor(f).astype(int); f=(f-s).astype(f32)
        lo=s<0; f[lo]=0; s[lo]=0
        hi=s>=n_src-1; f[hi]=0; s[hi]=n_src-1
        return s,f
    xs,fx=coeffs(W,w,sx); ys,fy=coeffs(H,h,sy)
    xs1=np.minimum(xs+1,w-1); ys1=np.minimum(ys+1,h-1)
    if src.ndim==3:
        fx_=fx[None,:,None]; fy_=fy[:,None,None]
    else:
        fx_=fx[None,:]; fy_=fy[:,None]
    rows=src[:,xs]*(f32(1)-fx_)+src[:,xs1]*fx_
    out=rows[ys]*(f32(1)-fy_)+rows[ys1]*fy_
    return out.astype(f32)
def gauss_kernel(n,sigma):
    # cv::getGaussianKernel(n, sigma, CV_32F)
    small={1:[1.],3:[0.25,0.5,0.25],5:[0.0625,0.25,0.375,0.25,0.0625],7:[0.03125,0.109375,0.21875,0.28125,0.21875,0.109375,0.03125]}
    if n<=7 and sigma<=0: return np.array(small[n],f32)
    sigmaX=sigma if sigma>0 else ((n-1)*0.5-1)*0.3+0.8
    scale2X=-0.5/(sigmaX*sigmaX)
    x=np.arange(n)-(n-1)*0.5
    k=np.exp(scale2X*x*x).astype(f32)   # stored as float then summed in double
    s=k.astype(np.float64).sum()
    return (k*f32(1.0/s)).astype(f32)
def gaussian_blur(img,ksize,sigma):
    k=gauss_kernel(ksize,sigma); r=ksize//2
    p=np.pad(img,((0,0),(r,r)),mode='reflect')
    W=img.shape[1]
    out=np.zeros_like(img)
    # symmetric accumulation: k0*c + sum k_i*(a+b)
    out=p[:,r:r+W]*k[r]
    for i in range(1,r+1): out=out+(p[:,r-i:r-i+W]+p[:,r+i:r+i+W])*k[r+i]
    p=np.pad(out,((r,r),(0,0)),mode='reflect'); H=img.shape[0]
    o2=p[r:r+H]*k[r]
    for i in range(1,r+1): o2=o2+(p[r-i:r-i+H]+p[r+i:r+i+H])*k[r+i]
    return o2.astype(f32)
def prepare_gaussian(n,sigma):
    if sigma<np.finfo(f32).eps: sigma=n*0.3
    x=np.arange(-n,n+1)
    g=np.exp(-x*x/(2*sigma*sigma)).astype(f32)
    s=1.0/g.astype(np.float64).sum()
    g=(g*s).astype(f32)   # (float)(g[x]*s) double mult
    g=(g.astype(np.float64)).astype(f32)
    xg=(x*g.astype(np.float64)).astype(f32); xxg=(x*x*g.astype(np.float64)).astype(f32)
    G=np.zeros((6,6))
    gd=g.astype(np.float64)
    for y in range(-n,n+1):
        for xx in range(-n,n+1):
            gg=float(f32(g[y+n]*g[xx+n]))  # g[y]*g[x] float mult? in C: G(0,0) += g[y]*g[x] : float*float -> float, added to double
            G[0,0]+=gg; G[1,1]+=float(f32(f32(gg)*f32(xx*xx))) if False else gg*xx*xx
            G[3,3]+=gg*xx*xx*xx*xx; G[5,5]+=gg*xx*xx*y*y
    G[2,2]=G[0,3]=G[0,4]=G[3,0]=G[4,0]=G[1,1]; G[4,4]=G[3,3]; G[3,4]=G[4,3]=G[5,5]
    invG=np.linalg.inv(G)
    return g,xg,xxg,invG[1,1],invG[0,3],invG[3,3],invG[5,5]
def poly_exp(src,n,sigma):
    H,W=src.shape
    g,xg,xxg,ig11,ig03,ig33,ig55=prepare_gaussian(n,sigma)
    c=n
    row0=src*g[c]; row1=np.zeros_like(src); row2=np.zeros_like(src)
    ys=np.arange(H)
    for k in range(1,n+1):
        s0=src[np.maximum(ys-k,0)]; s1=src[np.minimum(ys+k,H-1)]
        p=s0+s1
        row0=row0+g[c+k]*p; row1=row1+xg[c+k]*(s1-s0); row2=row2+xxg[c+k]*p
    def padx(a): return np.pad(a,((0,0),(n,n)),mode='edge').astype(np.float64)
    R0,R1,R2=padx(row0),padx(row1),padx(row2)
    xs=np.arange(W)+n
    b1=R0[:,xs]*float(g[c]); b2=np.zeros((H,W)); b3=R1[:,xs]*float(g[c]); b4=np.zeros((H,W)); b5=R2[:,xs]*float(g[c]); b6=np.zeros((H,W))
    for k in range(1,n+1):
        tg=(R0[:,xs+k].astype(f32)+R0[:,xs-k].astype(f32)).astype(np.float64)  # float add then to double
        b1+=tg*float(g[c+k]); b4+=tg*float(xxg[c+k])
        b2+=(R0[:,xs+k].astype(f32)-R0[:,xs-k].astype(f32)).astype(np.float64)*float(xg[c+k])
        b3+=(R1[:,xs+k].astype(f32)+R1[:,xs-k].astype(f32)).astype(np.float64)*float(g[c+k])
        b6+=(R1[:,xs+k].astype(f32)-R1[:,xs-k].astype(f32)).astype(np.float64)*float(xg[c+k])
        b5+=(R2[:,xs+k].astype(f32)+R2[:,xs-k].astype(f32)).astype(np.float64)*float(g[c+k])
    out=np.zeros((H,W,5),f32)
    out[...,1]=(b2*ig11); out[...,0]=(b3*ig11); out[...,3]=(b1*ig03+b4*ig33); out[...,2]=(b1*ig03+b5*ig33); out[...,4]=(b6*ig55)
    return out
BORDER=np.array([0.14,0.14,0.4472,0.4472,0.4472],f32)
def update_matrices(R0,R1,flow):
    H,W=flow.shape[:2]
    ys,xs=np.mgrid[0:H,0:W]
    dx=flow[...,0]; dy=flow[...,1]
    fx=(xs.astype(f32)+dx).astype(f32); fy=(ys.astype(f32)+dy).astype(f32)
    x1=np.floor(fx).astype(int); y1=np.floor(fy).astype(int)
    fx=(fx-x1.astype(f32)).astype(f32); fy=(fy-y1.astype(f32)).astype(f32)
    inside=(x1>=0)&(x1<W-1)&(y1>=0)&(y1<H-1)
    xc=np.clip(x1,0,W-2); yc=np.clip(y1,0,H-2)
    a00=(f32(1)-fx)*(f32(1)-fy); a01=fx*(f32(1)-fy); a10=(f32(1)-fx)*fy; a11=fx*fy
    def samp(c): return a00*R1[yc,xc,c]+a01*R1[yc,xc+1,c]+a10*R1[yc+1,xc,c]+a11*R1[yc+1,xc+1,c]
    r2=np.where(inside,samp(0),0).astype(f32); r3=np.where(inside,samp(1),0).astype(f32)
    r4=np.where(inside,(R0[...,2]+samp(2))*f32(0.5),R0[...,2]).astype(f32)
    r5=np.where(inside,(R0[...,3]+samp(3))*f32(0.5),R0[...,3]).astype(f32)
    r6=np.where(inside,(R0[...,4]+samp(4))*f32(0.25),R0[...,4]*f32(0.5)).astype(f32)
    r2=(R0[...,0]-r2)*f32(0.5); r3=(R0[...,1]-r3)*f32(0.5)
    r2=r2+r4*dy+r6*dx; r3=r3+r6*dy+r5*dx
    sc=np.ones((H,W),f32)
    bx=np.ones(W,f32); by=np.ones(H,f32)
    for i in range(min(5,W)):
        pass
    xi=np.arange(W); yi=np.arange(H)
    sx=np.where(xi<5,BORDER[np.minimum(xi,4)],f32(1))*np.where(xi>=W-5,BORDER[np.clip(W-xi-1,0,4)],f32(1))
    sy=np.where(yi<5,BORDER[np.minimum(yi,4)],f32(1))*np.where(yi>=H-5,BORDER[np.clip(H-yi-1,0,4)],f32(1))
    sc=(sx[None,:].astype(f32)*sy[:,None].astype(f32)).astype(f32)
    # order in C: ((sx_lo*sx_hi)*sy_lo)*sy_hi ; fine
    r2,r3,r4,r5,r6=[(v*sc).astype(f32) for v in (r2,r3,r4,r5,r6)]
    M=np.zeros((H,W,5),f32)
    M[...,0]=r4*r4+r6*r6; M[...,1]=(r4+r5)*r6; M[...,2]=r5*r5+r6*r6; M[...,3]=r4*r2+r6*r3; M[...,4]=r6*r2+r5*r3
    return M
def update_flow_blur(R0,R1,flow,M,block,update):
    H,W=flow.shape[:2]; m=block//2
    Md=M.astype(np.float64)
    # box window rows: y-m .. y+m with clamping, as implied by the running-sum construction
    # vsum(y) = sum_{j=y-m..y+m} M[clamp(j)]   (check init: vsum=M0*(m+2)+sum_{1..m-1} then += M[min(y+m)] - M[max(y-m-1,0)])
    ys=np.arange(H)
    vs=np.zeros_like(Md)
    for j in range(-m,m+1): vs+=Md[np.clip(ys+j,0,H-1)]
    xs=np.arange(W)
    hs=np.zeros_like(Md)
    for j in range(-m,m+1): hs+=vs[:,np.clip(xs+j,0,W-1)]
    scale=1.0/(block*block)
    g11,g12,g22,h1,h2=[hs[...,i]*scale for i in range(5)]
    idet=1.0/(g11*g22-g12*g12+1e-3)
    out=np.zeros_like(flow)
    out[...,0]=((g11*h2-g12*h1)*idet).astype(f32); out[...,1]=((g22*h1-g12*h2)*idet).astype(f32)
    Mn=update_matrices(R0,R1,out) if update else M
    return out,Mn
def farneback(prev,nxt,levels=10,pyr_scale=0.8,winsize=None,iters=7,poly_n=None,poly_sigma=None):
    H,W=prev.shape
    if poly_sigma is None: poly_sigma=(H+W)/1000.0
    if winsize is None: winsize=(H+W)//100
    if poly_n is None: poly_n=5 if poly_sigma<1.5 else 7
    min_size=32
    scale=1.0; k=0
    while k<levels:
        scale*=pyr_scale
        if W*scale<min_size or H*scale<min_size: break
        k+=1
    levels=k
    prev_flow=None
    for k in range(levels,-1,-1):
        scale=1.0
        for i in range(k): scale*=pyr_scale
        sigma=(1./scale-1)*0.5
        smooth_sz=int(round(sigma*5))|1   # cvRound
        smooth_sz=max(smooth_sz,3)
        w=int(round(W*scale)); h=int(round(H*scale))
        if prev_flow is None: flow=np.zeros((h,w,2),f32)
        else:
            flow=resize_linear(prev_flow,w,h)*f32(1./pyr_scale)
        R=[]
        for img in (prev,nxt):
            fimg=img.astype(f32)
            fimg=gaussian_blur(fimg,smooth_sz,sigma)
            I=resize_linear(fimg,w,h)
            R.append(poly_exp(I,poly_n,poly_sigma))
        M=update_matrices(R[0],R[1],flow)
        for i in range(iters):
            flow,M=update_flow_blur(R[0],R[1],flow,M,winsize,i<iters-1)
        prev_flow=flow
    return flow
