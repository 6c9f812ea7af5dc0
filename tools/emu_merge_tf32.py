#!/usr/bin/env python
"""CPU emulation of merge_tf32_kernel (ml-ease_b200/csrc/k3_cholesky.cu): the same tile / fragment / shared-memory index
expressions thread by thread in numpy, with the mma.sync.m16n8k8 fragment layout spelled out, run over all merge levels of a
small factor (partial last block, several tiles per block) and compared with numpy's inverse.  Written to check the index logic
of the kernel on a machine without a GPU; scratch tool, not part of the product or the tests."""
import numpy as np
TM=TN=128; TK=16; TA_LD=TK+4; TB_LD=TN+8
def tf32(x):
    f=np.float32(x); u=np.array(f,dtype=np.float32).view(np.uint32)
    u=(u+np.uint32(0x1000))&np.uint32(0xffffe000)   # rna to 10 bits mantissa (ties away), fine for emulation
    return u.view(np.float32)
def run_cta(mode,m,ldh,Lc,Yinv,Hinv,bx,by):
    r0=2*by*m; m2=min(m,ldh-r0-m)
    if m2<=0: return
    M,N=m2,m
    if mode==1:
        K=m; A=(Lc,(r0+m)*ldh+r0); B=(Yinv,r0*ldh+r0); C=(Hinv,(r0+m)*ldh+r0)
    else:
        K=m2; A=(Yinv,(r0+m)*ldh+(r0+m)); B=(Hinv,(r0+m)*ldh+r0); C=(Yinv,(r0+m)*ldh+r0)
    tiles_n=(N+TN-1)//TN
    i0=(bx//tiles_n)*TM; j0=(bx%tiles_n)*TN
    if i0>=M: return
    klo,khi=0,K
    if mode==1: klo=j0
    else: khi=min(K,i0+TM)
    As=np.zeros((2,TM*TA_LD),np.float32); Bs=np.zeros((2,TK*TB_LD),np.float32)
    acc=np.zeros((256,4,4,4),np.float32)
    tid=np.arange(256)
    pa=A[1]+(i0+(tid>>3))*ldh+klo+(tid&7)*2
    pbk=B[1]+(klo+(tid>>6))*ldh+j0+(tid&63)*2
    a_step=32*ldh; b_step=4*ldh; b_adv=TK*ldh
    ra=np.zeros((256,4,2)); rb=np.zeros((256,4,2))
    def gload():
        nonlocal pa,pbk
        for q in range(4):
            ok=(i0+(tid>>3)+32*q)<M
            idx=pa+q*a_step
            for t in range(256):
                ra[t,q]=A[0][idx[t]:idx[t]+2] if ok[t] else 0
            okb=(j0+(tid&63)*2)<N
            idx=pbk+q*b_step
            for t in range(256):
                rb[t,q]=B[0][idx[t]:idx[t]+2] if okb[t] else 0
        pa=pa+TK; pbk=pbk+b_adv
    def sstore(buf):
        for q in range(4):
            row=(tid>>3)+32*q; kk=(tid&7)*2
            As[buf][row*TA_LD+kk]=tf32(ra[:,q,0]); As[buf][row*TA_LD+kk+1]=tf32(ra[:,q,1])
            kk=(tid>>6)+4*q; jj=(tid&63)*2
            Bs[buf][kk*TB_LD+jj]=tf32(rb[:,q,0]); Bs[buf][kk*TB_LD+jj+1]=tf32(rb[:,q,1])
    warp=tid>>5; lane=tid&31; g=lane>>2; tg=lane&3
    wm=(warp&1)*64; wn=(warp>>1)*32
    def mma(accv,fa,fb):
        # accv [256,4], fa [256,4], fb [256,2]; per warp semantics of m16n8k8 row.col
        for w in range(8):
            sl=slice(32*w,32*w+32)
            Am=np.zeros((16,8),np.float32); Bm=np.zeros((8,8),np.float32); Cm=np.zeros((16,8),np.float32)
            for l in range(32):
                gg,tt=l>>2,l&3
                Am[gg,tt]=fa[32*w+l,0]; Am[gg+8,tt]=fa[32*w+l,1]; Am[gg,tt+4]=fa[32*w+l,2]; Am[gg+8,tt+4]=fa[32*w+l,3]
                Bm[tt,gg]=fb[32*w+l,0]; Bm[tt+4,gg]=fb[32*w+l,1]
                Cm[gg,2*tt]=accv[32*w+l,0]; Cm[gg,2*tt+1]=accv[32*w+l,1]; Cm[gg+8,2*tt]=accv[32*w+l,2]; Cm[gg+8,2*tt+1]=accv[32*w+l,3]
            Cm=Cm+Am@Bm
            for l in range(32):
                gg,tt=l>>2,l&3
                accv[32*w+l]=[Cm[gg,2*tt],Cm[gg,2*tt+1],Cm[gg+8,2*tt],Cm[gg+8,2*tt+1]]
    if klo<khi:
        gload(); sstore(0); buf=0
        k0=klo
        while k0<khi:
            more=k0+TK<khi
            if more: gload()
            a=As[buf]; b=Bs[buf]
            for k8 in range(0,TK,8):
                fa=np.zeros((4,256,4),np.float32); fb=np.zeros((4,256,2),np.float32)
                for f in range(4):
                    pa_=(wm+f*16+g)*TA_LD+k8+tg
                    fa[f,:,0]=a[pa_]; fa[f,:,1]=a[pa_+8*TA_LD]; fa[f,:,2]=a[pa_+4]; fa[f,:,3]=a[pa_+8*TA_LD+4]
                    pb_=(k8+tg)*TB_LD+wn+f*8+g
                    fb[f,:,0]=b[pb_]; fb[f,:,1]=b[pb_+4*TB_LD]
                for fm in range(4):
                    for fn in range(4):
                        mma(acc[:,fm,fn],fa[fm],fb[fn])
            if more: sstore(buf^1)
            buf^=1; k0+=TK
    sgn=1.0 if mode==1 else -1.0
    for fm in range(4):
        for h in range(2):
            i=i0+wm+fm*16+g+8*h
            for fn in range(4):
                j=j0+wn+fn*8+2*tg
                for t in range(256):
                    if i[t]>=M or j[t]>=N: continue
                    o=C[1]+i[t]*ldh+j[t]
                    C[0][o]=sgn*acc[t,fm,fn,2*h]; C[0][o+1]=sgn*acc[t,fm,fn,2*h+1]

def merges(ldh,L,leaf=256):
    Lc=L.copy().reshape(-1); Y=np.zeros(ldh*ldh); H=np.zeros(ldh*ldh)
    Y2=Y.reshape(ldh,ldh)
    for c in range(0,ldh,leaf):
        e=min(ldh,c+leaf); Y2[c:e,c:e]=np.linalg.inv(L[c:e,c:e])
    m=leaf
    while m<ldh:
        nmerge=(ldh+2*m-1)//(2*m); t=(m+TM-1)//TM
        for mode in (1,2):
            for by in range(nmerge):
                for bx in range(t*t):
                    run_cta(mode,m,ldh,Lc,Y,H,bx,by)
        m*=2
    return Y.reshape(ldh,ldh)
if __name__=="__main__":
    import sys
    rng=np.random.default_rng(0)
    for ldh in (352,608):
        G=rng.normal(size=(ldh,ldh))/np.sqrt(ldh); Hm=G@G.T+np.eye(ldh)
        L=np.linalg.cholesky(Hm)
        Y=merges(ldh,np.tril(L))
        ref=np.linalg.inv(L)
        print(ldh,"max err",np.abs(Y-ref).max(),"rel",np.abs(Y-ref).max()/np.abs(ref).max(),"upper max",np.abs(np.triu(Y,1)).max())
