"""Python emulation of the edge patch of k_blur7 (csrc/k_describe.cu): the PRMT selectors a lane computes for the
left / right image edge must reproduce BORDER_REFLECT_101 for every window byte a tap reads, for every width and
pitch the pyramid can produce. Pure logic check, no GPU."""
import numpy as np
def reflect101(p,n):
    if p<0: return -p
    if p>=n: return 2*(n-1)-p
    return p
def prmt(a,b,sel):
    by=[(a>>(8*i))&255 for i in range(4)]+[(b>>(8*i))&255 for i in range(4)]
    return sum(by[(sel>>(4*i))&7]<<(8*i) for i in range(4))
def check(w,pitch):
    rng=np.random.default_rng(w)
    row=rng.integers(0,256,pitch+8).astype(np.int64)  # beyond w: garbage
    for x in range(0,w,4):
        ld0=x>=4; ld2=x+8<=pitch
        def word(c): return int(sum(int(row[c+i])<<(8*i) for i in range(4)))
        W=[word(x-4) if ld0 else 0, word(x), word(x+4) if ld2 else 0]
        edge = x<4 or x+8>w
        sel=[0x3210,0x7654,0x7654]; hi=[False,False,True]
        if edge:
            for k in range(3):
                j=[-1]*4; lo=12; top=-1
                for b in range(4):
                    c=x-4+4*k+b
                    if -3<=c<=w+2:
                        j[b]=reflect101(c,w)-(x-4); lo=min(lo,j[b]); top=max(top,j[b])
                up= top>=8; base=4 if up else 0
                s=0
                for b in range(4):
                    v=(lo if lo<12 else 4*k+b) if j[b]<0 else j[b]
                    assert top<0 or 0<=v-base<8,(w,x,k,b,v,base)
                    s|=(v-base)<<(4*b)
                if top<0: sel[k]=0x3210 if k==0 else 0x7654; hi[k]= k==2
                else: sel[k]=s; hi[k]=up
            N=[prmt(W[1] if hi[k] else W[0], W[2] if hi[k] else W[1], sel[k]) for k in range(3)]
            W=N
        win=[(W[i//4]>>(8*(i%4)))&255 for i in range(12)]
        for k in range(4):
            if x+k>=w: continue
            for tap in range(-3,4):
                c=x+k+tap
                assert win[k+4+tap]==row[reflect101(c,w)],(w,x,k,tap)
def test_edge_selectors_reproduce_reflect101():
  for w in list(range(67,140))+[640,752,533,444,370,309,257,214,179,627,1280,1067]:
    for pitch in {w, (w+3)//4*4, (w+63)//64*64}:
        if pitch%4: continue
        check(w,pitch)
