"""CPU restatement of the DEVICE contour algorithm (csrc/contours.cu): shape algebra on label maps instead of polygons.
TEST INFRASTRUCTURE ONLY (tests/): it pins the identities the CUDA stage is built on to the reference's goldens without a
GPU, and is the thing to diff the kernels' intermediate results against when a GPU test fails.

The reference (scf/segmentation/base_cluster_based_dataset_segmenter.py:148-450,
scf/segmentation/black_white_handwritten_printed_text_segmenter.py:42-99) works on cv2 contours.  The same results follow from
  * filled contour  = 8-connected component of the complement of the outside background (`fill_outside` + `ndimage.label`),
  * contourArea     = pixels - L/2 - 1, L = boundary cracks - convex corners (`crack_stats`; Pick's theorem),
  * boundingRect    = min / max of the pixels,
  * merge_contours  = closure of "strict bounding-box test and a common pixel" under union + hole filling (`Seg.fixpoint`),
  * classification  = overlap sums with the strict bounding-box test, rendering = per-pixel lookup (`segment_one`).
Contour ORDER is not modelled: `segment_one` returns drop = None when the reference's drop rule (first contour of a class)
could depend on it, exactly the cases the device flags for the host path.  scipy.ndimage does the labelling here.
"""

import numpy as np, cv2
from scipy import ndimage as ndi
S8 = np.ones((3,3),int); S4 = ndi.generate_binary_structure(2,1)
CROSS = cv2.getStructuringElement(cv2.MORPH_CROSS,(3,3)).astype(np.uint8)

def fill_outside(U):
    H,W = U.shape
    P = np.zeros((H+2,W+2),bool); P[1:-1,1:-1] = U
    bg,_ = ndi.label(~P, structure=S4)
    return ~(bg == bg[0,0])[1:-1,1:-1]

def crack_stats(F):
    """Fc, Lc for a single filled binary shape F"""
    P = np.zeros((F.shape[0]+2,F.shape[1]+2),bool); P[1:-1,1:-1]=F
    O=~P
    up,dn,lf,rt = O[:-2,1:-1],O[2:,1:-1],O[1:-1,:-2],O[1:-1,2:]
    ul,ur,dl,dr = P[:-2,:-2],P[:-2,2:],P[2:,:-2],P[2:,2:]
    cr = up.astype(int)+dn+lf+rt
    cv = (up&lf&~ul).astype(int)+(up&rt&~ur)+(dn&lf&~dl)+(dn&rt&~dr)
    return int(F.sum()), int(((cr-cv)*F).sum())

def strict(a,b): return a[0]<b[2] and a[2]>b[0] and a[1]<b[3] and a[3]>b[1]

class Seg:
    """one (image, stage, class) segment: shapes from several keys, merged to a fixpoint"""
    def __init__(self, masks, merge=True):
        # masks: list over keys of uint8 [S,S]
        self.S = masks[0].shape[0]
        self.labs=[]; self.pix=[]   # shape pixel sets
        self.key_counts=[]
        for m in masks:
            D = cv2.morphologyEx(m, cv2.MORPH_DILATE, CROSS)
            F = fill_outside(D>0)
            lab,n = ndi.label(F, structure=S8)
            self.key_counts.append(n)
            base=len(self.pix)
            for i in range(1,n+1): self.pix.append(lab==i)
            self.labs.append(np.where(lab>0, lab-1+base, -1))
        n=len(self.pix)
        self.parent=list(range(n))
        self.fill={}   # root -> fill mask
        self.bbox=[self._bbox(p) for p in self.pix]
        if merge and len(masks)>1: self.fixpoint()
    def _bbox(self,m):
        ys,xs=np.nonzero(m); return (xs.min(),ys.min(),xs.max(),ys.max())
    def find(self,i):
        while self.parent[i]!=i: i=self.parent[i]
        return i
    def groups(self):
        g={}
        for i in range(len(self.pix)): g.setdefault(self.find(i),[]).append(i)
        return g
    def gmask(self,root,members):
        U=np.zeros((self.S,self.S),bool)
        for m in members: U|=self.pix[m]
        return U
    def fixpoint(self):
        n=len(self.pix)
        gb={i:self.bbox[i] for i in range(n)}
        cover_fill=np.full((self.S,self.S),-1)
        while True:
            changed=False
            # inner: pair detection
            while True:
                ch=False
                roots=[np.where(l>=0, np.array([self.find(i) for i in range(n)]+[-1])[l], -1) for l in self.labs]
                fr=np.where(cover_fill>=0, np.array([self.find(i) for i in range(n)]+[-1])[cover_fill], -1)
                layers=roots+[fr]
                pairs=set()
                for a in range(len(layers)):
                    for b in range(a+1,len(layers)):
                        m=(layers[a]>=0)&(layers[b]>=0)&(layers[a]!=layers[b])
                        if m.any():
                            pa=np.stack([layers[a][m],layers[b][m]],1)
                            for x,y in np.unique(pa,axis=0): pairs.add((int(x),int(y)))
                for x,y in pairs:
                    rx,ry=self.find(x),self.find(y)
                    if rx!=ry and strict(gb[rx],gb[ry]):
                        r=min(rx,ry); o=max(rx,ry); self.parent[o]=r
                        bx,by=gb[rx],gb[ry]
                        gb[r]=(min(bx[0],by[0]),min(bx[1],by[1]),max(bx[2],by[2]),max(bx[3],by[3]))
                        ch=True
                if not ch: break
                changed=True
            if not changed: break
            # fills for merged groups
            for root,mem in self.groups().items():
                if len(mem)>1:
                    U=self.gmask(root,mem)
                    F=fill_outside(U)
                    fl=F&~U
                    self.fill[root]=fl
                    cover_fill[fl]=root
        self.gb=gb
    def final(self):
        """list of dict(root, members, mask F, Fc, Lc, bbox)"""
        out=[]
        for root,mem in self.groups().items():
            U=self.gmask(root,mem)
            F=U|self.fill.get(root,False) if len(mem)>1 else U
            fc,lc=crack_stats(F)
            out.append(dict(root=root,n=len(mem),F=F,Fc=fc,Lc=lc,bbox=self._bbox(F),area=fc-lc/2-1))
        return out

def segment_one(masks_by_key, det_keys, fine_keys, class_names, colors, image_size, only_keep_overlapping, min_area, fine_class='printed_text'):
    """masks_by_key[key][class] -> uint8 [S,S].  Returns (uint8 [S,S,3] label image, drop) with drop True / False, or None when
    the reference's decision depends on contour order (the device flags those images for the host path)."""
    names=[n for n in class_names if n!='background']
    regions={}
    for c in names:
        seg=Seg([masks_by_key[k][c] for k in det_keys])
        valid=all(n>0 for n in seg.key_counts)
        fin=seg.final() if valid else []
        if len(det_keys)>1:
            fin=[g for g in fin if (g['n']>1 or not only_keep_overlapping)]
        fin=[g for g in fin if g['area']>=min_area]
        regions[c]=fin if fin else None
    seg=Seg([masks_by_key[k][fine_class] for k in fine_keys])
    valid=all(n>0 for n in seg.key_counts)
    fine=seg.final() if valid else []
    if len(fine_keys)>1: fine=[g for g in fine if g['n']>1]
    live=[c for c in names if regions[c] is not None]
    picked={c:[] for c in names}
    for f in fine:
        best,bs=None,0
        for c in live:
            sc=0
            for r in regions[c]:
                if strict(f['bbox'],r['bbox']): sc+=int((f['F']&r['F']).sum())
            if sc>bs: best,bs=c,sc
        if best is not None: picked[best].append(f)
    for c in names: picked[c]=[f for f in picked[c] if f['area']>=min_area]
    # determine_images_to_drop reads the FIRST contour of each class: decided only when every contour of the class agrees
    limit=int(image_size*0.95)
    certain=undecided=False
    for c in names:
        huge=sum(1 for f in picked[c] if (f['bbox'][2]-f['bbox'][0]+1)>limit and (f['bbox'][3]-f['bbox'][1]+1)>limit)
        certain|=huge>0 and huge==len(picked[c])
        undecided|=0<huge<len(picked[c])
    drop=True if certain else (None if undecided else False)
    ink=masks_by_key[fine_keys[-1]][fine_class]>0
    out=np.empty((image_size,image_size,3),np.uint8); out[:]=np.asarray(colors['background'],np.uint8)
    for c in names:
        for f in picked[c]:
            out[f['F']&ink]=np.asarray(colors[c],np.uint8)
    return out, drop
