"""CPU stand-in for libaoenv_b200.so used ONLY by the `-m "not gpu"` tests to exercise the Python host layer
(object protocol, step sequencing, buffer bookkeeping) without a GPU.  It implements the C ABI of
include/aoenv.h on raw host addresses with numpy, following the oracle's arithmetic.  It is test
infrastructure: the product never imports it and has no CPU path."""
import ctypes as C

import numpy as np
import torch



def _arr(ptr, shape, dtype=np.float32):
    if ptr is None:
        return None
    addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
    if addr is None or addr == 0:
        return None
    n = int(np.prod(shape))
    ct = {np.float32: C.c_float, np.int32: C.c_int32, np.uint8: C.c_uint8, np.float64: C.c_double, np.int64: C.c_int64}[dtype]
    return np.ctypeslib.as_array((ct * n).from_address(addr)).reshape(shape)


def _val(x):
    return x.value if hasattr(x, "value") else x


def _f2o(f):
    i = np.float32(f).view(np.int32)
    return i if i >= 0 else np.int32(i ^ np.int32(0x7fffffff))


def _o2f(i):
    i = np.int32(i)
    return (i if i >= 0 else np.int32(i ^ np.int32(0x7fffffff))).view(np.float32)


class FakeLib:
    def __init__(self):
        self.launches = 0
        self.rs = np.random.RandomState(123)

    def aoenv_abi_version(self):
        return 1

    def aoenv_last_error(self):
        return b""

    def aoenv_detector_integrate(self, frame, B, rows, cols, det, stream):
        """OOPAO/Detector.py:190-301 with numpy's generators, restricted to the stages in det.reserved (0 = all)."""
        self.launches += 1
        d = det._obj if hasattr(det, "_obj") else det
        st = d.reserved or 7
        f = _arr(frame, (B, rows, cols))
        x = f.astype(np.float64)
        if st & 1:
            if d.photon_noise:
                x = self.rs.poisson(np.maximum(x, 0)).astype(np.float64)
            x = x * d.qe
        if st & 2:
            if d.dark_electrons > 0:
                x = x + self.rs.poisson(d.dark_electrons, size=x.shape)
            if d.has_fwc:
                x = np.clip(x, 0, d.fwc)
            if d.sensor_emccd:
                x = x * d.gain
        if st & 4:
            if d.readout_noise:
                x = x + np.round(self.rs.normal(size=x.shape) * d.readout_noise)
            if not d.sensor_emccd:
                x = x * d.gain
            if d.bits > 0:
                full = float((1 << d.bits) - 1)
                x = np.minimum(np.trunc(x / d.fwc * full), full)
        f[:] = x.astype(np.float32)
        return 0

    def aoenv_psf_image(self, opd_a, opd_b, pupil, amp, w_planes, B, R, N, os_, win, phase_scale, field_planes, ldk, work_t,
                        planes_u, work_f, psf, psf_max, stream):
        """OOPAO/Telescope.py:296-360 with numpy's FFT: padded field, half-pixel phasor, centred FFT / N, |.|^2, central
        crop, oversampling binned away."""
        self.launches += 5
        a, b_ = _arr(opd_a, (B, R, R)), _arr(opd_b, (B, R, R))
        pu, am = _arr(pupil, (R, R)).astype(np.float64), _arr(amp, (R, R)).astype(np.float64)
        out, mx = _arr(psf, (B, win, win)), _arr(psf_max, (B,))
        pad = (N - R) // 2
        k = np.arange(N)
        xx, yy = np.meshgrid(k, k)
        phasor = np.exp(-1j * np.pi / N * (xx + yy))
        size = os_ * win
        lo = N // 2 - size // 2
        for e in range(B):
            t = a[e].astype(np.float64) + (b_[e] if b_ is not None else 0)
            field = np.pad(am * np.exp(1j * t * pu * float(_val(phase_scale))), pad)
            emf = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(field * phasor)) / N)[lo:lo + size, lo:lo + size]
            img = (np.abs(emf) ** 2).reshape(win, os_, win, os_).sum(axis=(1, 3))
            out[e] = img.astype(np.float32)
            mx[e] = out[e].max()
        return 0

    def aoenv_set_wfs6_variant(self, factorised):
        return 0

    def aoenv_launch_count(self):
        return self.launches

    def aoenv_set_pdl(self, enabled):
        return 1

    def aoenv_sh_step(self, c, part, opd_a, rows_cur, det, action, coefs_next, rows_next, obs, reward, strehl, total, residual,
                      stream):
        """The six calls of csrc/step.cu, in its order (parts: bit 0 the first two, bit 1 the next two, bit 2 the last two)."""
        from rlao_b200 import _lib as L
        c = c._obj if hasattr(c, "_obj") else c
        if part & 1:
            dm = L.DmSepStruct()
            dm.rows, dm.wlr, dm.ilr, dm.nActP, dm.WL = rows_cur, c.dm.wlr, c.dm.ilr, c.dm.nActP, c.dm.WL
            self.aoenv_shwfs_frame_dm(opd_a, None, dm, c.order, c.pupil, c.amp, c.valid, c.B, c.nS, c.n, c.phase_scale, det, 0,
                                      c.frame, c.envmax, c.stats, stream)
            self.aoenv_shwfs_slopes(c.frame, c.envmax, 0, c.valid_idx, c.nV, c.ref_xy, c.inv_units, c.threshold_cog, c.B, c.nS,
                                    c.n, c.slopes, c.lds, None, 2, stream)
        if part & 2:
            self.aoenv_gemm_tn(c.slopes, c.lds, c.rec_f32, c.lds, c.rec, c.ldr, c.B, c.nA, c.lds, 1.0, stream)
            self.aoenv_observe(c.rec, c.ldr, c.act_idx, c.B, c.nA, c.nAct2, c.stats, c.n_pupil, c.phase_scale, obs, reward,
                               strehl, total, residual, stream)
        if part & 4:
            self.aoenv_command_update(action, c.act_idx, c.B, c.nA, c.nAct2, c.leak, coefs_next, c.dm_prev, c.ldc, stream)
            self.aoenv_dm_rows(coefs_next, c.ldc, c.act_pos, c.nA, c.nAct, c.dm.nActP, c.wx, c.j0x, c.W, c.B, c.nS * c.n,
                               rows_next, stream)
        return 0

    def aoenv_atm_update(self, state, w_planes, opd_out, stream):
        raise AssertionError("the CPU stand-in sequences frames in Python (AOENV_ATM_NATIVE=0)")

    # ---- atmosphere (sliding-window canvas) ---------------------------------------------------------------
    @staticmethod
    def _window(ptr, B, M, pitch, env_stride):
        """[B, M, M] strided view of a window given the address of its origin pixel."""
        addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
        n = (B - 1) * env_stride + (M - 1) * pitch + M
        flat = np.ctypeslib.as_array((C.c_float * n).from_address(addr))
        return np.lib.stride_tricks.as_strided(flat, shape=(B, M, M), strides=(4 * env_stride, 4 * pitch, 4))

    @staticmethod
    def _key(v):
        i = np.asarray(v, dtype=np.float32).view(np.int32).astype(np.int64)
        o = np.where(i >= 0, i, i ^ 0x7fffffff)
        return (o & 0xffffffff) ^ 0x80000000

    @staticmethod
    def _unkey(packed):
        k = ((int(packed) >> 32) & 0xffffffff) ^ 0x80000000
        k = k - (1 << 32) if k >= (1 << 31) else k
        i = np.int32(k)
        return (i if i >= 0 else np.int32(i ^ np.int32(0x7fffffff))).view(np.float32)

    def aoenv_atm_gather(self, win, B, M, pitch, env_stride, sx, sy, inner_rc, nI, nO, xi, seed, stream_id, zx, ldz,
                         zx_planes, parts, stream):
        self.launches += 1
        addr = (win.value if isinstance(win, C.c_void_p) else int(win)) - 4 * (sy * pitch + sx)
        m = self._window(addr, B, M, pitch, env_stride)       # window after the shift (old content)
        rc = _arr(inner_rc, (nI, 2), np.int32)
        z = _arr(zx, (B, ldz))
        z[:] = 0
        z[:, :nI] = m[:, rc[:, 0], rc[:, 1]]
        x = _arr(xi, (B, nO))
        if x is None:       # counter-based like the kernel: the draw depends on (seed, stream id) only
            x = np.random.RandomState((int(_val(seed)) * 1000003 + int(_val(stream_id))) % (2 ** 32)).normal(size=(B, nO))
        z[:, nI:nI + nO] = x
        return 0

    def aoenv_atm_ring(self, win, B, M, pitch, env_stride, win_offset, nO, X, ldx, ext, flag, force_rescan, stream):
        self.launches += 2
        m = self._window(win, B, M, pitch, env_stride)
        x = _arr(X, (B, ldx))
        e = _arr(ext, (B, 2), np.int64)
        outer = np.ones((M, M), dtype=bool)
        outer[1:-1, 1:-1] = False
        for b in range(B):
            m[b][outer] = x[b, :nO]
            flatpos = win_offset + np.arange(M)[:, None] * pitch + np.arange(M)[None, :]
            keys = (self._key(m[b]).astype(np.uint64) << np.uint64(32)) | flatpos.astype(np.uint64)
            e[b, 0] = np.int64(np.uint64(keys.min()).astype(np.int64))
            e[b, 1] = np.int64(np.uint64(keys.max()).astype(np.int64))
        return 0

    def aoenv_atm_gather_multi(self, wins, sx, sy, seeds, stream_ids, G, B, M, pitch, env_stride, inner_rc, nI, nO, xi, zx, ldz,
                               zx_planes, parts, stream):
        z0 = zx.value if isinstance(zx, C.c_void_p) else int(zx)
        for g in range(G):
            self.aoenv_atm_gather(int(wins[g]), B, M, pitch, env_stride, int(sx[g]), int(sy[g]), inner_rc, nI, nO, None,
                                  int(seeds[g]), int(stream_ids[g]), z0 + 4 * g * B * ldz, ldz, None, parts, stream)
        self.launches -= G - 1
        return 0

    def aoenv_atm_ring_multi(self, wins, win_offsets, exts, G, B, M, pitch, env_stride, nO, X, ldx, flag, force_rescan, stream):
        x0 = X.value if isinstance(X, C.c_void_p) else int(X)
        for g in range(G):
            self.aoenv_atm_ring(int(wins[g]), B, M, pitch, env_stride, int(win_offsets[g]), nO, x0 + 4 * g * B * ldx, ldx,
                                int(exts[g]), flag, force_rescan, stream)
        self.launches -= 2 * (G - 1)
        return 0

    def aoenv_atm_compact(self, src, dst, B, M, pitch, env_stride, ext, pos_delta, stream):
        self.launches += 1
        s_, d_ = self._window(src, B, M, pitch, env_stride), self._window(dst, B, M, pitch, env_stride)
        d_[:] = s_
        e = _arr(ext, (B, 2), np.int64)
        u = e.view(np.uint64)
        u[:] = (u & np.uint64(0xffffffff00000000)) | ((u & np.uint64(0xffffffff)) + np.uint64(pos_delta % (1 << 32))) & np.uint64(0xffffffff)
        return 0

    def aoenv_vk_screens(self, seed, screen0, S, N, amp, inject, wa, wb, Kp, parts, sh_ex, sh_ey, h_amp, h_mx, h_my, wp, wa_, lda,
                         wb_, ldb, dst, pitch, env_stride, stream):
        """numpy FFT of the same spectrum (the kernels run it as two DFT-matrix products); sub-harmonics from the tables."""
        self.launches += 5
        a = _arr(amp, (N, N)).astype(np.float64)                    # sqrt(PSD) del_f (-1)^(a+b)
        k = np.arange(N)
        sign = 1.0 - 2.0 * ((k[:, None] + k[None, :]) % 2)
        inj = _arr(inject, (S, 2, N, N))
        ex = _arr(sh_ex, (3, 2, N, 2)).astype(np.float64)
        ex = ex[..., 0] + 1j * ex[..., 1]
        ey = _arr(sh_ey, (3, 2, N, 2)).astype(np.float64)
        ey = ey[..., 0] + 1j * ey[..., 1]
        la = _arr(h_amp, (3, 2, 2))
        out = self._window(dst, S, N, pitch, env_stride)
        for s_ in range(S):
            if inj is not None:
                re, im = inj[s_, 0].astype(np.float64), inj[s_, 1].astype(np.float64)
            else:
                rs = np.random.RandomState((int(_val(seed)) * 7919 + int(_val(screen0)) + s_) % (2 ** 32))
                re, im = rs.normal(size=(N, N)), rs.normal(size=(N, N))
            cn = (re + 1j * im) * a * sign                          # amp carries the fftshift signs: undo them
            hi = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(cn))).real
            flat = re.reshape(-1)
            lo = np.zeros((N, N), dtype=complex)
            for p in range(3):
                for i in range(2):
                    for j in range(2):
                        c = (flat[18 * p + 3 * i + j] + 1j * flat[18 * p + 9 + 3 * i + j]) * la[p, i, j]
                        lo += c * ey[p, i][:, None] * ex[p, j][None, :]
            out[s_] = (hi + lo.real - lo.real.mean()).astype(np.float32)
        return 0

    def aoenv_atm_phase(self, h_canvas, h_ext, h_org, L, B, R, M, Mc, pitch, fp_off, roff, coff, wr, wc, wt, opd_scale,
                        opd_out, stream):
        self.launches += 1
        env_stride = Mc * pitch
        h_map = [int(h_canvas[l]) + 4 * (h_org[2 * l] * pitch + h_org[2 * l + 1]) for l in range(L)]
        out = _arr(opd_out, (B, R, R))
        acc = np.zeros((B, R, R), dtype=np.float32)
        for l in range(L):
            m = self._window(h_map[l], B, M, pitch, env_stride)
            e = _arr(h_ext[l], (B, 2), np.int64).view(np.uint64)
            for b in range(B):
                v = np.zeros((R, R), dtype=np.float32)
                for pr in range(4):
                    h = np.zeros((R, R), dtype=np.float32)
                    r0 = fp_off + roff[l] + pr
                    for pc in range(4):
                        c0 = fp_off + coff[l] + pc
                        h += np.float32(wc[4 * l + pc]) * m[b, r0:r0 + R, c0:c0 + R]
                    v += np.float32(wr[4 * l + pr]) * h
                v = np.clip(v, self._unkey(e[b, 0]), self._unkey(e[b, 1]))
                acc[b] += np.float32(wt[l]) * v
        out[:] = acc * np.float32(_val(opd_scale))
        return 0

    # ---- gemm ------------------------------------------------------------------------------------------
    def aoenv_split_bf16(self, src, lds, rows, K, parts, dst, ldk, stream):
        self.launches += 1          # operand planes are never read by this stand-in (its GEMMs take the float32 operands)
        return 0

    def aoenv_gemm_tn(self, X, ldx, W, ldw, D, ldd, M, N, K, alpha, stream):
        self.launches += 1
        assert K % 16 == 0
        x, w, d = _arr(X, (M, ldx)), _arr(W, (N, ldw)), _arr(D, (M, ldd))
        d[:, :N] = (np.float32(_val(alpha)) * (x[:, :K].astype(np.float64) @ w[:, :K].astype(np.float64).T)).astype(np.float32)
        return 0

    def aoenv_dm_surface_separable(self, coefs, ldc, act_pos, nA, nAct, gx, gy, band_x, band_y, wx, j0x, wyp, i0y, W, B, R,
                                   opd, stream):
        self.launches += 1
        c = _arr(coefs, (B, ldc))
        pos = _arr(act_pos, (nA,), np.int32)
        gx_, gy_ = _arr(gx, (nAct, R)).astype(np.float64), _arr(gy, (nAct, R)).astype(np.float64)
        bx, by = _arr(band_x, (R, 2), np.int32), _arr(band_y, (R, 2), np.int32)
        j = np.arange(nAct)[:, None]
        mx = (j >= bx[None, :, 0]) & (j <= bx[None, :, 1])
        my = (j >= by[None, :, 0]) & (j <= by[None, :, 1])
        out = _arr(opd, (B, R, R))
        for b in range(B):
            C_ = np.zeros(nAct * nAct)
            C_[pos] = c[b, :nA]
            out[b] = ((gy_ * my).T @ C_.reshape(nAct, nAct) @ (gx_ * mx)).astype(np.float32)
        return 0

    # ---- WFS -------------------------------------------------------------------------------------------
    def aoenv_shwfs_frame(self, opd_a, opd_b, pupil, amp, valid, B, nS, n, phase_scale, det, shared_max, frame, envmax,
                          stats, stream):
        self.launches += 2
        R = nS * n
        a, b_ = _arr(opd_a, (B, R, R)), _arr(opd_b, (B, R, R))
        pu, am = _arr(pupil, (R, R)), _arr(amp, (R, R))
        va = _arr(valid, (nS * nS,), np.uint8).astype(bool)
        fr = _arr(frame, (B, R, R))
        em = _arr(envmax, (1 if shared_max else B,), np.int32)
        st = _arr(stats, (B, 4), np.float64)
        scale = np.float32(_val(phase_scale))
        N = 2 * n
        k = np.arange(N)
        xx, yy = np.meshgrid(k, k)
        phasor = np.exp(-(1j * np.pi * (N + 1) / N) * (xx + yy))
        tiles = lambda img: img.reshape(nS, n, nS, n).transpose(0, 2, 3, 1).reshape(nS * nS, n, n)
        lo = N // 2 - n // 2
        maxes = []
        for e in range(B):
            t = a[e] + (b_[e] if b_ is not None else 0)
            if st is not None:
                m = pu > 0
                da = (a[e] - a[e][R // 2, R // 2])[m].astype(np.float64)
                dt = (t - t[R // 2, R // 2])[m].astype(np.float64)
                st[e] = [da.sum(), (da ** 2).sum(), dt.sum(), (dt ** 2).sum()]
            ph = (t * pu * scale).astype(np.float64)
            field = np.zeros((nS * nS, N, N), dtype=complex)
            field[:, lo:lo + n, lo:lo + n] = np.exp(1j * tiles(ph)) * tiles(am.astype(np.float64))
            I = np.abs(np.fft.fft2(field * phasor, axes=(1, 2)) / N) ** 2
            spots = I.reshape(-1, n, 2, n, 2).sum(axis=(2, 4))
            spots[~va] = 0
            f = spots.reshape(nS, nS, n, n).transpose(0, 2, 1, 3).reshape(R, R)
            assert det is None or _val(det) is None or not det, "fake backend: detector chain not modelled"
            fr[e] = f.astype(np.float32)
            maxes.append(fr[e].reshape(nS, n, nS, n).transpose(0, 2, 1, 3).reshape(nS * nS, n, n)[va].max())
        if shared_max:
            em[0] = _f2o(max(maxes))
        else:
            for e in range(B):
                em[e] = _f2o(maxes[e])
        return 0

    def aoenv_shwfs_fused_smem(self, nS, n, rows_max, groups, t_rows, WL):
        if rows_max <= 0 or rows_max > nS:
            return -1
        R, rows_px = nS * n, rows_max * n
        up = lambda v: (v + 127) & ~127
        wstride = 2 * (((max(WL, 14) + 1) // 2 + 3) // 4 * 4)
        o = up(rows_px * R * 4)
        o = up(o + rows_px * R)
        o = up(o + groups * n * n * 256)
        if t_rows > 0:
            o = up(o + t_rows * R * 4)
            o = up(o + rows_px * wstride * 4)
        return up(o + 2048)

    def aoenv_shwfs_frame_dm(self, opd_a, opd_b, dm, order, pupil, amp, valid, B, nS, n, phase_scale, det, shared_max, frame,
                             envmax, stats, stream):
        """The DM surface from its factored form (T rows + row-weight windows), then the ordinary frame (`order` only
        changes which thread works on which lenslet)."""
        R = nS * n
        if order:
            o = _arr(order, (nS * nS,), np.int32)
            va = _arr(valid, (nS * nS,), np.uint8).astype(bool)
            assert sorted(o.tolist()) == list(range(nS * nS)) and va[o[:va.sum()]].all()
        d = dm._obj if hasattr(dm, "_obj") else dm
        if d is None or not getattr(d, "rows", None):
            return self.aoenv_shwfs_frame(opd_a, opd_b, pupil, amp, valid, B, nS, n, phase_scale, det, shared_max, frame,
                                          envmax, stats, stream)
        assert not opd_b
        WL, half = d.WL, (d.WL + 1) // 2
        hp = (half + 3) // 4 * 4
        trows = _arr(d.rows, (B, d.nActP, R)).astype(np.float64)
        wlr, ilr = _arr(d.wlr, (R, 2 * hp)).astype(np.float64), _arr(d.ilr, (nS,), np.int32)
        surf = np.zeros((B, R, R), dtype=np.float32)
        for y in range(R):
            i0 = int(ilr[y // n])
            w = np.array([wlr[y, (t // half) * hp + t % half] for t in range(WL)])
            surf[:, y, :] = np.einsum("t,btx->bx", w, trows[:, i0:i0 + WL, :])
        self._inline_dm_surface = surf                    # kept alive for the duration of the call below
        return self.aoenv_shwfs_frame(opd_a, C.c_void_p(surf.ctypes.data), pupil, amp, valid, B, nS, n, phase_scale, det,
                                      shared_max, frame, envmax, stats, stream)

    def aoenv_dm_rows(self, coefs, ldc, act_pos, nA, nAct, nActP, wx, j0x, W, B, R, rows, stream):
        self.launches += 1
        c = _arr(coefs, (B, ldc))
        pos = _arr(act_pos, (nA,), np.int32)
        w, j0 = _arr(wx, (R, W)).astype(np.float64), _arr(j0x, (R,), np.int32)
        out = _arr(rows, (B, nActP, R))
        for b in range(B):
            Cimg = np.zeros((nAct, nAct + W))
            Cimg[pos // nAct, pos % nAct] = c[b, :nA]
            for x in range(R):
                out[b, :nAct, x] = Cimg[:, j0[x]:j0[x] + W] @ w[x]
        return 0

    def aoenv_shwfs_fused(self, opd_a, opd_b, dm, pupil8, amp0, row_start, order, nlit, slot_of, B, nS, n, cluster, groups,
                          phase_scale, ref_xy, nV, inv_units, thr, frame, slopes, lds, slope_planes, parts, envmax, stats, stream):
        """Strip by strip, like the kernel: the DM surface of a strip from the window tables and the rows of T the strip
        may read ([ilr[first row], ilr[first row] + t_rows)), lenslets visited through `order` / `nlit`, slopes scattered
        by `slot_of`."""
        self.launches += 1
        R = nS * n
        rs = [int(row_start[k]) for k in range(cluster + 1)]
        assert rs[0] == 0 and rs[-1] == nS
        rows_max = max(rs[k + 1] - rs[k] for k in range(cluster))
        a = _arr(opd_a, (B, R, R))
        b_ = _arr(opd_b, (B, R, R))
        pu = _arr(pupil8, (R, R), np.uint8).astype(bool)
        od, nl = _arr(order, (cluster, rows_max * nS), np.int32), _arr(nlit, (cluster,), np.int32)
        so = _arr(slot_of, (nS * nS,), np.int32)
        fr = _arr(frame, (B, R, R))
        sl = _arr(slopes, (B, lds)) if slopes else None
        ref = _arr(ref_xy, (2, nV)) if sl is not None else None
        em = _arr(envmax, (B,), np.int32)
        st = _arr(stats, (B, 4), np.float64)
        scale, a0 = np.float32(_val(phase_scale)), float(_val(amp0))
        d = dm._obj if hasattr(dm, "_obj") else dm
        sep = d is not None and d.rows
        if sep:
            WL, half = d.WL, (d.WL + 1) // 2
            hp = (half + 3) // 4 * 4
            trows = _arr(d.rows, (B, d.nActP, R)).astype(np.float64)
            wlr, ilr = _arr(d.wlr, (R, 2 * hp)).astype(np.float64), _arr(d.ilr, (nS,), np.int32)
        N = 2 * n
        k = np.arange(N)
        xx, yy = np.meshgrid(k, k)
        phasor = np.exp(-(1j * np.pi * (N + 1) / N) * (xx + yy))
        lo = N // 2 - n // 2
        for e in range(B):
            total = a[e].astype(np.float32).copy()
            if b_ is not None:
                total = total + b_[e]
            spots_img = np.zeros((R, R), dtype=np.float64)
            lit_max = -np.inf
            for r in range(cluster):
                y0, y1 = rs[r] * n, rs[r + 1] * n
                if sep:
                    tBase = int(ilr[rs[r]])
                    assert int(ilr[rs[r + 1] - 1]) + WL - tBase <= d.t_rows, "t_rows too small for this strip"
                    assert tBase + d.t_rows <= d.nActP + WL
                    for y in range(y0, y1):
                        i0 = int(ilr[y // n])
                        w = np.array([wlr[y, (t // half) * hp + t % half] for t in range(WL)])
                        total[y] = total[y] + (w @ trows[e, i0:i0 + WL]).astype(np.float32)
                ph = (total[y0:y1] * np.float32(scale)).astype(np.float64)
                field_px = np.where(pu[y0:y1], a0 * np.exp(1j * ph), 0)
                for i in range((rs[r + 1] - rs[r]) * nS):
                    lens = int(od[r, i])
                    lr, l = lens // nS, lens % nS
                    if i >= nl[r]:
                        continue
                    tile = field_px[lr * n:(lr + 1) * n, l * n:(l + 1) * n].T
                    fld = np.zeros((N, N), dtype=complex)
                    fld[lo:lo + n, lo:lo + n] = tile
                    I = np.abs(np.fft.fft2(fld * phasor) / N) ** 2
                    sp = I.reshape(n, 2, n, 2).sum(axis=(1, 3))
                    spots_img[y0 + lr * n:y0 + (lr + 1) * n, l * n:(l + 1) * n] = sp
                    lit_max = max(lit_max, np.float32(sp.max()))
            if st is not None:
                av, tv = a[e][pu].astype(np.float64), total[pu].astype(np.float64)
                st[e] = [av.sum(), (av ** 2).sum(), tv.sum(), (tv ** 2).sum()]
            if fr is not None:
                fr[e] = spots_img.astype(np.float32)
            if em is not None:
                em[e] = _f2o(lit_max if sl is not None else -np.inf)
            if sl is not None:
                f32 = spots_img.astype(np.float32).astype(np.float64)
                for kk in np.nonzero(so >= 0)[0]:
                    li, lj = kk // nS, kk % nS
                    im = f32[li * n:(li + 1) * n, lj * n:(lj + 1) * n].copy()
                    im[im < np.float32(_val(thr)) * np.float32(lit_max)] = 0
                    with np.errstate(invalid="ignore", divide="ignore"):
                        s_ = im.sum()
                        cx = (im * np.arange(n)[:, None]).sum() / s_
                        cy = (im * np.arange(n)[None, :]).sum() / s_
                    cx = cx if np.isfinite(cx) else 0.0
                    cy = cy if np.isfinite(cy) else 0.0
                    t = so[kk]
                    sl[e, t] = (cx - ref[0, t]) * np.float32(_val(inv_units))
                    sl[e, nV + t] = (cy - ref[1, t]) * np.float32(_val(inv_units))
        return 0

    def aoenv_shwfs_camera(self, *a):
        raise NotImplementedError("fake backend: detector chain not modelled")

    def aoenv_shwfs_slopes(self, frame, envmax, shared_max, valid_idx, nV, ref_xy, inv_units, thr, B, nS, n, slopes, lds,
                           slope_planes, parts, stream):
        self.launches += 1
        R = nS * n
        fr = _arr(frame, (B, R, R))
        em = _arr(envmax, (1 if shared_max else B,), np.int32)
        vi = _arr(valid_idx, (nV,), np.int32)
        ref = _arr(ref_xy, (2, nV))
        sl = _arr(slopes, (B, lds))
        for e in range(B):
            maps = fr[e].reshape(nS, n, nS, n).transpose(0, 2, 1, 3).reshape(nS * nS, n, n)[vi].astype(np.float64)
            mx = _o2f(em[0 if shared_max else e])
            # ShackHartmannOracle.centroid thresholds at threshold * maps.max(); feed the kernel's max explicitly
            im = maps.copy()
            im[im < np.float32(_val(thr)) * mx] = 0
            with np.errstate(invalid="ignore", divide="ignore"):
                s = im.sum(axis=(1, 2))
                cx = (im * np.arange(n)[None, :, None]).sum(axis=(1, 2)) / s
                cy = (im * np.arange(n)[None, None, :]).sum(axis=(1, 2)) / s
            cx[~np.isfinite(cx)] = 0
            cy[~np.isfinite(cy)] = 0
            sl[e, :nV] = (cx - ref[0]) * np.float32(_val(inv_units))
            sl[e, nV:2 * nV] = (cy - ref[1]) * np.float32(_val(inv_units))
        return 0

    def aoenv_shwfs_measure_f64(self, opd, pupil, amp, valid, valid_idx, nV, ref_xy, inv_units, thr, F, nS, n, phase_scale,
                                shared_max, frame, envmax, slopes, lds, stream):
        self.launches += 2
        R = nS * n
        a = _arr(opd, (F, R, R))
        pu, am = _arr(pupil, (R, R)), _arr(amp, (R, R))
        va = _arr(valid, (nS * nS,), np.uint8).astype(bool)
        vi = _arr(valid_idx, (nV,), np.int32)
        ref = _arr(ref_xy, (2, nV), np.float64)
        fr = _arr(frame, (F, R, R), np.float64)
        sl = _arr(slopes, (F, lds), np.float64)
        N = 2 * n
        k = np.arange(N)
        xx, yy = np.meshgrid(k, k)
        phasor = np.exp(-(1j * np.pi * (N + 1) / N) * (xx + yy))
        tiles = lambda img: img.reshape(nS, n, nS, n).transpose(0, 2, 3, 1).reshape(nS * nS, n, n)
        lo = N // 2 - n // 2
        maps = []
        for e in range(F):
            ph = a[e].astype(np.float64) * pu * float(_val(phase_scale))
            field = np.zeros((nS * nS, N, N), dtype=complex)
            field[:, lo:lo + n, lo:lo + n] = np.exp(1j * tiles(ph)) * tiles(am.astype(np.float64))
            I = np.abs(np.fft.fft2(field * phasor, axes=(1, 2)) / N) ** 2
            spots = I.reshape(-1, n, 2, n, 2).sum(axis=(2, 4))
            spots[~va] = 0
            fr[e] = spots.reshape(nS, nS, n, n).transpose(0, 2, 1, 3).reshape(R, R)
            maps.append(spots[vi])
        gmax = max(m.max() for m in maps)
        for e in range(F):
            im = maps[e].copy()
            im[im < float(_val(thr)) * (gmax if shared_max else im.max())] = 0
            with np.errstate(invalid="ignore", divide="ignore"):
                s_ = im.sum(axis=(1, 2))
                cx = (im * np.arange(n)[None, :, None]).sum(axis=(1, 2)) / s_
                cy = (im * np.arange(n)[None, None, :]).sum(axis=(1, 2)) / s_
            cx[~np.isfinite(cx)] = 0
            cy[~np.isfinite(cy)] = 0
            sl[e, :nV] = (cx - ref[0]) * float(_val(inv_units))
            sl[e, nV:2 * nV] = (cy - ref[1]) * float(_val(inv_units))
        return 0

    # ---- control ---------------------------------------------------------------------------------------
    def aoenv_command_update(self, action, act_idx, B, nA, nAct2, leak, coefs, dm_prev, ldc, stream):
        self.launches += 1
        a = _arr(action, (B, nAct2))
        idx = _arr(act_idx, (nA,), np.int32)
        c, p = _arr(coefs, (B, ldc)), _arr(dm_prev, (B, ldc))
        new = p[:, :nA] * np.float32(_val(leak)) + a[:, idx] * np.float32(1e-6)
        c[:, :nA] = new
        p[:, :nA] = new
        return 0

    def aoenv_normal_fill(self, seed, counter, rows, cols, ld, sigma, out, planes, parts, stream):
        self.launches += 1
        o = _arr(out, (rows, ld))
        o[:] = 0
        rs = np.random.RandomState((int(_val(seed)) * 1000003 + int(_val(counter))) % (2 ** 32))
        o[:, :cols] = rs.normal(size=(rows, cols)) * np.float32(_val(sigma))
        return 0

    def aoenv_vec_to_img(self, vec, ldv, act_idx, B, nA, nAct2, scale, img, stream):
        self.launches += 1
        v, idx, o = _arr(vec, (B, ldv)), _arr(act_idx, (nA,), np.int32), _arr(img, (B, nAct2))
        o[:] = 0
        o[:, idx] = v[:, :nA] * np.float32(_val(scale))
        return 0

    def aoenv_observe(self, rec, ldr, act_idx, B, nA, nAct2, stats, n_pupil, phase_scale, obs, reward, strehl, total,
                      residual, stream):
        self.launches += 1
        r = _arr(rec, (B, ldr))
        idx = _arr(act_idx, (nA,), np.int32)
        o = _arr(obs, (B, nAct2))
        st = _arr(stats, (B, 4), np.float64)
        o[:] = 0
        o[:, idx] = -r[:, :nA] * np.float32(1e6)
        _arr(reward, (B,))[:] = -np.sqrt((o.astype(np.float64) ** 2).sum(axis=1))
        if st is not None:
            n = _val(n_pupil)
            va = np.maximum(st[:, 1] / n - (st[:, 0] / n) ** 2, 0)
            vt = np.maximum(st[:, 3] / n - (st[:, 2] / n) ** 2, 0)
            _arr(total, (B,))[:] = np.sqrt(va) * 1e9
            _arr(residual, (B,))[:] = np.sqrt(vt) * 1e9
            _arr(strehl, (B,))[:] = np.exp(-vt * float(_val(phase_scale)) ** 2)
        return 0

    def aoenv_pyramid_supported(self, N):
        return int(N in (128, 288))

    def aoenv_pyramid_frames(self, opd_a, opd_b, pupil, amp, lin, mod, mask_s, B, R, N, nTheta, bin_, phase_scale, wx1, wyt,
                             intensity, frame, stream):
        """numpy FFTs of the same chain (OOPAO/Pyramid.py:469-504,581-603,987-1002); the mask arrives in the kernels'
        digit-scrambled order and is put back in natural order here."""
        self.launches += 4
        a, b_ = _arr(opd_a, (B, R, R)), _arr(opd_b, (B, R, R))
        pu, am = _arr(pupil, (R, R)).astype(np.float64), _arr(amp, (R, R)).astype(np.float64)
        ln = _arr(lin, (R,)).astype(np.float64)
        md = _arr(mod, (nTheta, 2)).astype(np.float64)
        ms = _arr(mask_s, (N, N, 2)).astype(np.float64)
        ms = ms[..., 0] + 1j * ms[..., 1]
        N1, N2 = 16, N // 16
        pos = np.arange(N)
        freq = (pos // N2) + N1 * (pos % N2)                         # frequency held by each scrambled position
        mask = np.zeros((N, N), dtype=complex)
        mask[np.ix_(freq, freq)] = ms
        k = np.arange(N)
        ph1 = np.exp(-1j * np.pi * (N + 1) / N * k)
        phasor = ph1[:, None] * ph1[None, :]
        lo = (N - R) // 2
        out, inten = _arr(frame, (B, N // bin_, N // bin_)), _arr(intensity, (B, N, N))
        for e in range(B):
            t = a[e].astype(np.float64) + (b_[e] if b_ is not None else 0)
            acc = np.zeros((N, N))
            for th in range(nTheta):
                pm = (md[th, 0] * ln[None, :] + md[th, 1] * ln[:, None]) * pu
                sup = np.zeros((N, N), dtype=complex)
                sup[lo:lo + R, lo:lo + R] = am * np.exp(1j * (t * pu * float(_val(phase_scale)) + pm))
                acc += np.abs(np.fft.ifft2(np.fft.fft2(sup * phasor) * mask)) ** 2
            inten[e] = acc.astype(np.float32)
            nc = N // bin_
            out[e] = acc.reshape(nc, bin_, nc, bin_).sum(axis=(1, 3)).astype(np.float32)
        return 0

    def aoenv_psf_peak(self, *a):
        raise NotImplementedError("fake backend: psf_peak")


def install(monkeypatch):
    """Routes rlao_b200._lib to the fake backend and lets objects live on the CPU."""
    from rlao_b200 import _lib, gemm
    fake = FakeLib()
    monkeypatch.setattr(gemm, "BACKEND", "simt")
    monkeypatch.setenv("AOENV_ATM_NATIVE", "0")      # frames sequenced by the Python layer (aoenv_atm_update launches kernels)
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda device: torch.device("cpu"))
    monkeypatch.setattr(_lib, "stream_ptr", lambda device=None: None)
    monkeypatch.setattr(_lib, "launch_count", lambda: fake.launches)
    return fake
