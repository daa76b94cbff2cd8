"""A SECOND, independent restatement of the reference's MATLAB "normal version" (main.m and its helpers), written
line by line from the .m text on top of numpy / scipy / torch primitives.  TEST INFRASTRUCTURE (like oracle/): it
exists so that oracle/cds_oracle.c — the C restatement every GPU parity test is compared with — has a second
opinion for every Image Processing Toolbox primitive whose semantics are otherwise taken from its documentation:

    imfilter(A, h)                 -> scipy.ndimage.correlate(mode='constant') with the kernel centre of imfilter
    fspecial('gaussian', k, s)     -> the published formula incl. the eps*max truncation, as a full 2-D kernel
    imresize(A, [r c], 'bilinear') -> torch.nn.functional.interpolate(mode='bilinear', antialias=True) away from the
                                      border (torch clips the window at the border, MATLAB mirrors it) and
                                      imresize.m's contributions() for the whole map
    medfilt2(A, [2 2])             -> brute-force 2x2 window (ordfilt2's domain centre floor((size+1)/2), zero padding,
                                      mean of the two middle values as medfilt2.m does for an even element count)
    round                          -> half away from zero

Nothing here is imported by the product package.  Each function cites the reference lines it follows."""
import numpy as np

EPS = np.finfo(np.float64).eps


def m_round(x):
    """MATLAB round(): half away from zero."""
    x = np.asarray(x, np.float64)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


# ------------------------------------------------------------------------------------------------ small helpers
def upsample(T, multiply, orisize):
    """upsample.m:1-10; orisize = [w4 h4] as main.m:79 passes video_size/4."""
    T = np.asarray(T, np.float64)
    m, n = T.shape
    out = np.zeros((int(orisize[1]), int(orisize[0])))
    for ri in range(1, m + 1):
        for ci in range(1, n + 1):
            r1 = min(ri * multiply, orisize[1]); c1 = min(ci * multiply, orisize[0])
            out[(ri - 1) * multiply:r1, (ci - 1) * multiply:c1] = T[ri - 1, ci - 1]
    return out


def fourX2(T):
    """fourX2.m:1-17 — every element becomes a 4x4 block."""
    return np.kron(np.asarray(T, np.float64), np.ones((4, 4)))


def pm_norm(a):
    """pm_norm.m:1-5 (0/0 -> NaN on a constant map, as MATLAB)."""
    a = np.asarray(a, np.float64)
    t = a - a.min()
    with np.errstate(invalid="ignore", divide="ignore"):
        return t / t.max()


def mysterythrehold(a, im_height, im_width):
    """mysterythrehold.m:3-7."""
    a = np.array(a, np.float64)
    ave = a.sum() / (im_height * im_width)
    big = a[a > (a.max() - ave)]
    with np.errstate(invalid="ignore", divide="ignore"):
        thresh = big.sum() / big.size if big.size else np.nan
    a[a > thresh] = thresh
    return a / (a.max() + 0.000001)


def immove(im, mov_x, mov_y):
    """immove.m:1-35, literally: pad, then slice with [m2, n2] = size(temp) (NOT temp2) — the reference's quirk."""
    im = np.asarray(im, np.float64)
    mov_x = int(mov_x); mov_y = int(mov_y)
    m, n = im.shape
    padX = np.zeros((m, abs(mov_x)))
    temp = np.hstack([padX, im]) if mov_x > 0 else np.hstack([im, padX])
    padY = np.zeros((abs(mov_y), n + abs(mov_x)))
    temp2 = np.vstack([padY, temp]) if mov_y > 0 else np.vstack([temp, padY])
    m2, n2 = temp.shape
    if mov_x > 0:
        if mov_y > 0:
            return temp2[0:m, 0:n]
        return temp2[m2 - m:m2, 0:n]
    if mov_y > 0:
        return temp2[0:m, n2 - n:n2]
    return temp2[m2 - m:m2, n2 - n:n2]


def getdistMatrix(hei, wid):
    """getdistMatrix.m:1-17."""
    x = np.arange(1, hei + 1)[:, None]; y = np.arange(1, wid + 1)[None, :]
    d = np.floor(np.sqrt((x - hei // 2) ** 2.0 + (y - wid // 2) ** 2.0))
    return 1 - d / d.max()


# ------------------------------------------------------------------------------------------------ IPT primitives
def fspecial_gaussian(hsize, sigma):
    """fspecial('gaussian', hsize, sigma) with a scalar hsize: the full 2-D kernel including h(h < eps*max) = 0."""
    siz = (hsize - 1) / 2.0
    v = np.arange(-siz, siz + 1e-9, 1.0)
    x, y = np.meshgrid(v, v)
    h = np.exp(-(x * x + y * y) / (2.0 * sigma * sigma))
    h[h < EPS * h.max()] = 0
    s = h.sum()
    return h / s if s != 0 else h


def imfilter(a, h):
    """imfilter(A, h): correlation, zero padding, 'same'; the kernel centre is floor((size(h)+1)/2) (1-based).
    scipy centres a length-k axis at k//2 (0-based), one further for even k: origin = -1 moves it back."""
    from scipy import ndimage
    origin = tuple(-1 if (k % 2 == 0) else 0 for k in h.shape)
    return ndimage.correlate(np.asarray(a, np.float64), h, mode="constant", cval=0.0, origin=origin)


def imfilter_bruteforce(a, h):
    """The definition itself (for small maps): out(i,j) = sum_uv h(u,v) a(i+u-cu, j+v-cv), zero outside."""
    a = np.asarray(a, np.float64)
    kr, kc = h.shape
    cu = (kr + 1) // 2 - 1; cv = (kc + 1) // 2 - 1
    p = np.zeros((a.shape[0] + kr, a.shape[1] + kc))
    p[cu:cu + a.shape[0], cv:cv + a.shape[1]] = a
    out = np.zeros_like(a)
    for u in range(kr):
        for v in range(kc):
            out += h[u, v] * p[u:u + a.shape[0], v:v + a.shape[1]]
    return out


def _contributions(in_length, out_length):
    """imresize.m contributions() for the 'bilinear' (triangle) kernel, antialiasing on (the default)."""
    scale = out_length / in_length
    kernel_width = 2.0
    tri = lambda t: np.where((t >= -1) & (t < 0), t + 1, np.where((t >= 0) & (t <= 1), 1 - t, 0.0))
    if scale < 1:
        h = lambda t: scale * tri(scale * t)
        kernel_width = kernel_width / scale
    else:
        h = tri
    x = np.arange(1, out_length + 1, dtype=np.float64)
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kernel_width / 2)
    P = int(np.ceil(kernel_width)) + 2
    indices = left[:, None] + np.arange(P)[None, :]
    weights = h(u[:, None] - indices)
    weights = weights / weights.sum(axis=1, keepdims=True)
    aux = np.concatenate([np.arange(1, in_length + 1), np.arange(in_length, 0, -1)])
    indices = aux[np.mod(indices.astype(np.int64) - 1, aux.size)]
    return weights, indices - 1      # 0-based


def imresize_bilinear(a, out_rows, out_cols):
    """imresize(A, [out_rows out_cols], 'bilinear'): separable, the dimension with the smaller scale first."""
    a = np.asarray(a, np.float64)
    dims = [(0, out_rows), (1, out_cols)]
    dims.sort(key=lambda d: d[1] / a.shape[d[0]])
    for axis, n_out in dims:
        w, idx = _contributions(a.shape[axis], n_out)
        a = np.moveaxis(a, axis, 0)
        a = np.einsum("op,op...->o...", w, a[idx])
        a = np.moveaxis(a, 0, axis)
    return a


def imresize_torch(a, out_rows, out_cols):
    """torch's antialiased bilinear resize: the same triangle kernel and pixel-centre convention as imresize; it differs
    only where the (widened) window crosses the border (clip + renormalise instead of MATLAB's mirror)."""
    import torch
    t = torch.from_numpy(np.asarray(a, np.float64))[None, None]
    r = torch.nn.functional.interpolate(t, size=(out_rows, out_cols), mode="bilinear", antialias=True, align_corners=False)
    return r[0, 0].numpy()


def medfilt2_2x2(a):
    """medfilt2(A, [2 2]): ordfilt2's window for an even domain is anchored at floor((size+1)/2) = (1,1), i.e. it covers
    rows i..i+1 and columns j..j+1; zeros beyond the border; 4 elements -> 0.5*(2nd) + 0.5*(3rd) smallest."""
    a = np.asarray(a, np.float64)
    p = np.zeros((a.shape[0] + 1, a.shape[1] + 1)); p[:-1, :-1] = a
    out = np.zeros_like(a)
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            v = np.sort(p[i:i + 2, j:j + 2].ravel())
            out[i, j] = 0.5 * v[1] + 0.5 * v[2]
    return out


# ------------------------------------------------------------------------------------------------ stage (c)
def direction_cluster(mv_x, mv_y, k=16):
    """direction_cluster.m:1-50 (mv_vote.m:3-4 calls it with k = 16)."""
    mv_x = np.asarray(mv_x, np.float64).ravel(); mv_y = np.asarray(mv_y, np.float64).ravel()
    ta = np.arctan2(mv_y, mv_x) / np.pi + 1
    ta = np.floor(ta / (2.0 / k))
    ta[ta == 16] = 15
    num = np.array([np.count_nonzero(ta == ii - 1) for ii in range(1, k + 1)])
    addr_num = np.flatnonzero(num == num.max())
    addr = ta == addr_num[0]
    xv = mv_x[addr]; yv = mv_y[addr]
    Len = xv.size; Lencut = Len // 2
    xs = np.sort(xv); ys = np.sort(yv)

    def half_mean(s):
        if abs(s[0]) > abs(s[Len - 1]):
            return s[Lencut - 1:Len].sum() / (Len - Lencut + 1)      # sort(Lencut:Len), 1-based
        return s[0:Lencut].sum() / Lencut
    return half_mean(xs), half_mean(ys)


def cm_judge(eVx, eVy, n_eff):
    """cm_judge.m:1-12 -> 1 when the camera moves."""
    V = eVx ** 2 + eVy ** 2
    return 0 if np.count_nonzero(V <= 2) > 0.5 * n_eff else 1


def camera_motion(mvx, mvy, bitmap, depth, w, h):
    """main.m:77-105 for one picture -> (Vx, Vy, mv_norm, bit_norm, depth_norm, x_cammov, y_cammov)."""
    Vx = np.asarray(mvx, np.float64) / 4; Vy = np.asarray(mvy, np.float64) / 4
    Bm = upsample(bitmap, 16, (w // 4, h // 4))
    eff = Bm < Bm.min() + 20
    Vxf = medfilt2_2x2(Vx); Vyf = medfilt2_2x2(Vy)
    eVx = Vxf.T[eff.T]; eVy = Vyf.T[eff.T]                 # find() runs column-major
    if cm_judge(eVx, eVy, eVx.size) == 1:
        x_ave, y_ave = direction_cluster(eVx, eVy)
        Vx = Vxf - x_ave; Vy = Vyf - y_ave
    else:
        x_ave = y_ave = 0.0
    mv_new = np.sqrt(Vx ** 2 + Vy ** 2)
    depth = np.asarray(depth, np.float64)
    return (Vx, Vy, mv_new / (mv_new.max() + EPS), Bm / (Bm.max() + EPS), depth / (depth.max() + EPS),
            int(m_round(x_ave)), int(m_round(y_ave)))


# ------------------------------------------------------------------------------------------------ stage (b)
def computecontrast5(pad, W, ln, win, m, n):
    """computecontrast5.cpp:52-96 on the zero-padded map (window rows r..r+win-1 of the padded map, centre at +len)."""
    pad = np.asarray(pad, np.float64)
    C = pad[ln:ln + m, ln:ln + n]
    out = np.zeros((m, n))
    for i in range(win):
        for j in range(win):
            p = pad[i:i + m, j:j + n]
            out += W[i, j] * np.where(p != 0, np.abs(p - C), 0.0)
    return out


def sudoku_contrast(sub, delta, cusize, patchsize):
    """Sudoku_contrastcpp5.m:1-52."""
    sub = np.asarray(sub, np.float64)
    m, n = sub.shape
    win = int(m_round(patchsize / cusize)); ln = win // 2
    X, Y = np.meshgrid(np.arange(1, win + 1), np.arange(1, win + 1))
    R = ((X - (ln + 1)) ** 2 + (Y - (ln + 1)) ** 2) ** 0.5
    W = np.exp(-R ** 2 / (2 * delta ** 2))
    p = pm_norm(np.pad(sub + 0.00000001, ln))
    c = computecontrast5(p, W, ln, win, m, n)
    with np.errstate(invalid="ignore", divide="ignore"):
        return c / c.max()


def cu_contrast5(salmap, cusize, delta, patchsize):
    """cu_contrast5.m:3-63."""
    salmap = np.asarray(salmap, np.float64)
    cusize = int(cusize)
    if cusize == 1:
        o = sudoku_contrast(salmap, delta, cusize, patchsize)
        with np.errstate(invalid="ignore", divide="ignore"):
            return o / o.max()
    M, N = salmap.shape
    nr = -(-M // cusize); nc = -(-N // cusize)
    sub = np.zeros((nr, nc))
    for k in range(nr):
        for j in range(nc):
            t = salmap[k * cusize:min((k + 1) * cusize, M) if k < nr - 1 else M, j * cusize:min((j + 1) * cusize, N) if j < nc - 1 else N]
            sub[k, j] = t.mean(axis=0).mean()            # mean(mean(temp)): column means, then their mean
    c = sudoku_contrast(sub, delta, cusize, patchsize)
    res = np.zeros((M, N))
    for k in range(nr):
        for j in range(nc):
            res[k * cusize:(k + 1) * cusize if k < nr - 1 else M, j * cusize:(j + 1) * cusize if j < nc - 1 else N] = c[k, j]
    with np.errstate(invalid="ignore", divide="ignore"):
        return res / res.max()


# ------------------------------------------------------------------------------------------------ stage (d)
def svm_rbf_decision(SV, coef, rho, gamma, X):
    """svm.cpp:2501-2575 for a two-class C-SVC with an RBF kernel (svm.cpp:325-365)."""
    d2 = ((X[:, None, :] - SV[None, :, :]) ** 2).sum(-1)
    return np.exp(-gamma * d2) @ coef - rho


# ------------------------------------------------------------------------------------------------ main.m:77-305
def process_clip(w, h, fps, mvx, mvy, depth, bitmap, model, mean_vec, std_vec, use_rbf=1, want_feat=False):
    """The normal version for a whole clip; mvx / mvy / depth [F][h4][w4], bitmap [F][rows of CTUs][cols of CTUs].
    -> (maps [F][h][w], cam [F][2], nine features per picture: at 200x200 (RBF branch) or H x W (linear branch))."""
    F = mvx.shape[0]
    SV, coef, rho, gamma, _ = model
    FeatureWeight = np.array([0.058, 0.171, -0.095, 0.068, -0.01, 0.509, -0.051, 0.154, 0.196])
    PS = int(m_round(fps * 0.3)); PT = int(m_round(PS * 7))
    d_mv_t, d_bit_t, d_dep_t = 26.0, 46.0, 46.0
    gk = int(m_round(h / 7.5)); gs = float(m_round(h / 30))
    G = fspecial_gaussian(gk, gs)
    dist = getdistMatrix(h, w)
    mv_x, mv_y, mv_norm, bit_norm, depth_norm = [], [], [], [], []
    cam = np.zeros((F, 2), np.int32)
    for ii in range(F):
        Vx, Vy, mn, bn, dn, cx, cy = camera_motion(mvx[ii], mvy[ii], bitmap[ii], depth[ii], w, h)
        mv_x.append(Vx); mv_y.append(Vy); mv_norm.append(mn); bit_norm.append(bn); depth_norm.append(dn)
        cam[ii] = (cx, cy)
    x_cammov, y_cammov = cam[:, 0], cam[:, 1]
    maps = np.zeros((F, h, w)); feats = []
    sm = {}
    nan0 = lambda a: np.where(np.isnan(a), 0.0, a)
    for k in range(1, F + 1):                                  # 1-based like main.m:137
        acc = [0.0] * 5; count = 0
        for ii in range(k - PS + 1, k + 1):
            if ii > 0:
                for q, src in enumerate((mv_norm, mv_x, mv_y, bit_norm, depth_norm)):
                    acc[q] = acc[q] + src[ii - 1]
                count += 1
        mv_crop, mvx_crop, mvy_crop, bit_crop, depth_crop = (a / count for a in acc)
        sm[k] = (bit_crop, depth_crop, mvx_crop, mvy_crop)
        mv_crop = mysterythrehold(mv_crop, h, w); bit_crop = mysterythrehold(bit_crop, h, w); depth_crop = mysterythrehold(depth_crop, h, w)
        cx = cy = 0
        mvsum = np.zeros((h // 4, w // 4)); bitsum = np.zeros_like(mvsum); depsum = np.zeros_like(mvsum)
        refbit, refdepth, refmvx, refmvy = sm[k]
        for j in range(1, PT):
            if k - j > 0:
                tb, td, tx, ty = sm[k - j]
                cx = x_cammov[k - j] + cx; cy = y_cammov[k - j] + cy          # x_cammov(k-j+1), 1-based
                sx, sy = m_round(cx / 4), m_round(cy / 4)
                tb, td, tx, ty = (immove(t, sx, sy) for t in (tb, td, tx, ty))
                mvsum = mvsum + np.exp(-j / d_mv_t) * np.sqrt((refmvx - tx) ** 2 + (refmvy - ty) ** 2)
                bitsum = bitsum + np.exp(-j / d_bit_t) * np.abs(refbit - tb)
                depsum = depsum + np.exp(-j / d_dep_t) * np.abs(refdepth - td)
        flt = lambda a: pm_norm(imfilter(fourX2(a), G))
        mv_ori, bit_ori, depth_ori = flt(mv_crop), flt(bit_crop), flt(depth_crop)
        mv_temporal, bit_temporal, depth_temporal = flt(mvsum), flt(bitsum), flt(depsum)
        bit_spacial = flt(cu_contrast5(bit_crop, 64 / 4, 3, 64 * 5 / 4))
        depth_spacial = flt(cu_contrast5(depth_crop, 8 / 4, 13, 8 * 24 / 4))
        mxc = nan0(cu_contrast5(mvx_crop, 1, 11, 4 * 48 / 4)); myc = nan0(cu_contrast5(mvy_crop, 1, 11, 4 * 48 / 4))
        mv_spacial = pm_norm(imfilter(np.sqrt(fourX2(mxc) ** 2 + fourX2(myc) ** 2), G))
        mv_ori, bit_ori, depth_ori, mv_temporal, bit_temporal, depth_temporal, bit_spacial, depth_spacial, mv_spacial = (
            nan0(a) for a in (mv_ori, bit_ori, depth_ori, mv_temporal, bit_temporal, depth_temporal, bit_spacial, depth_spacial, mv_spacial))
        depth_temporal = mysterythrehold(depth_temporal, h, w); bit_temporal = mysterythrehold(bit_temporal, h, w)
        mv_temporal = mysterythrehold(mv_temporal, h, w)
        nine = [mv_ori, mv_temporal, mv_spacial, bit_ori, bit_temporal, bit_spacial, depth_ori, depth_temporal, depth_spacial]
        if use_rbf:
            nine = [imresize_bilinear(a, 200, 200) for a in nine]
            X = np.stack([a.ravel(order="F") for a in nine], axis=1)
            X = (X - mean_vec[None, :]) / std_vec[None, :]
            dec = svm_rbf_decision(np.asarray(SV, np.float64), np.asarray(coef, np.float64), rho, gamma, X)
            comb = pm_norm(dec.reshape((200, 200), order="F"))
            comb = imfilter(imresize_bilinear(comb, h, w), G)
        else:
            X = np.stack([a.ravel(order="F") for a in nine], axis=1)
            X = (X - mean_vec[None, :]) / std_vec[None, :]
            comb = (X @ FeatureWeight).reshape((h, w), order="F")
        maps[k - 1] = pm_norm(comb * dist)
        if want_feat:
            feats.append(np.stack(nine))
    return maps, cam, (np.stack(feats) if want_feat else None)
