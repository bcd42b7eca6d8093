"""Static plan of the DRFNet family on the tap-GEMM primitive (reference: drf_net.py:23-147).

Everything here is host-side integer bookkeeping, done once per (architecture, dtype):
  * the phase-blocked layout of high-resolution maps (phase_table),
  * for every convolution of the net: the tap table of its forward pass, of its data-gradient
    and (same as forward) of its weight-gradient,
  * index maps from the flat fp32 parameter bucket (reference layouts: Conv2d [Cout,Cin,kh,kw],
    ConvTranspose2d [Cin,Cout,kh,kw]) to the packed weight slabs the kernels read, and back from
    packed weight-gradient slabs to the flat gradient bucket.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .ops import TapTable

# (kernel, stride, padding) of the projection pair — drf_net.py:70-77
PROJ = {2: (6, 2, 2), 3: (7, 3, 2), 4: (8, 4, 2), 8: (12, 8, 2)}
MAX_NT = 256


def phase_table(r: int) -> List[Tuple[int, int]]:
    """(py, px) of every phase slot of an r-times up-scaled map.  Power-of-two r: nested (Z-order)
    so that each nn.PixelShuffle(2) of the output block (drf_net.py:141-142) is a reinterpretation
    and the sub-pixel groups of the 8x8-stride-4 transposed convolution are contiguous."""
    if r in (1, 2, 4, 8):
        slots = [(0, 0)]
        rr = r
        while rr > 1:
            slots = [(2 * py + i, 2 * px + j) for (py, px) in slots for i in (0, 1) for j in (0, 1)]
            rr //= 2
        return slots
    return [(i, j) for i in range(r) for j in range(r)]


@dataclass
class ParamRef:
    name: str
    offset: int
    shape: Tuple[int, ...]

    def idx(self, *ix):
        """flat index (numpy broadcast) of element ix of this parameter."""
        flat = 0
        for d, i in zip(self.shape, ix):
            flat = flat * d + i
        return self.offset + flat


@dataclass
class Layer:
    """One tap-GEMM with its packed weights (slabs[t] is an int64 [nt,kc] array of flat parameter
    indices, -1 = structural zero) and optional packed bias (int64 [out_c] of flat indices)."""
    name: str
    table: TapTable
    slabs: List[np.ndarray]
    out_c: int
    bias_idx: Optional[np.ndarray] = None
    w_off: int = -1      # element offset into the packed weight buffer (set by Plan.finalize)
    b_off: int = -1      # element offset into the packed bias buffer
    slope: Optional[ParamRef] = None   # PReLU applied to this layer's output
    bias_c: int = 0      # number of distinct bias values (the packed bias has this period)

    @property
    def w_numel(self):
        return len(self.slabs) * self.table.nt * self.table.kc


def _kc(F: int, bf16: bool) -> int:
    if F % 64 == 0:
        return 64
    if bf16:
        raise ValueError(f"bf16/tcgen05 mode needs num_features % 64 == 0 (got {F}); use precision='fp32'")
    return F


def _split_nt(total: int) -> List[Tuple[int, int]]:
    """split `total` output channels into (o0, nt) runs with nt <= MAX_NT, all equal."""
    if total <= MAX_NT:
        return [(0, total)]
    n = -(-total // MAX_NT)
    while total % n:
        n += 1
    return [(i * (total // n), total // n) for i in range(n)]


class DrfPlan:
    """variant 'drf': DRFNet / DRFSISRNet (drf_net.py); 'srfb': SRFBNet (srfb_net.py:38-50,137-151) — the
    same feedback block, input block named `lrf_block`, reconstruction block `r_block` = transposed
    conv + PReLU + 3x3 conv, no feature skip (the global skip is the bilinear up-sampled input)."""

    def __init__(self, in_channels, out_channels, F, G, r, bf16, variant="drf"):
        self.variant = variant
        self.in_name = "in_block" if variant == "drf" else "lrf_block"
        if r not in PROJ:
            raise ValueError(f"The upscale factor should be 2, 3, 4 or 8. Got {r}.")
        if G + 1 > 8:
            raise ValueError("num_groups > 7 is not supported by the fused concat (8 sources per tap-GEMM)")
        self.cin, self.cout, self.F, self.G, self.r, self.bf16 = in_channels, out_channels, F, G, r, bf16
        self.kc = _kc(F, bf16)
        self.kb = F // self.kc
        self.k, self.s, self.p = PROJ[r]
        self.phases = phase_table(r)
        self.slot_of = {yx: i for i, yx in enumerate(self.phases)}
        self.params: Dict[str, ParamRef] = {}
        self.n_params = 0
        self.fwd: Dict[str, Layer] = {}
        self.bwd: Dict[str, Layer] = {}
        self._declare_params()
        self._build_layers()
        self._finalize()

    # ---- parameters in reference order (state_dict / named_parameters order of drf_net.py) ----
    def _add_param(self, name, shape):
        self.params[name] = ParamRef(name, self.n_params, tuple(shape))
        self.n_params += int(np.prod(shape))

    def _declare_params(self):
        F, G, k = self.F, self.G, self.k
        conv = lambda p, o, i, ks: (self._add_param(p + ".weight", (o, i, ks, ks)), self._add_param(p + ".bias", (o,)))
        dconv = conv  # ConvTranspose2d weight is [Cin, Cout, k, k] — same rank, roles swapped
        prelu = lambda p: self._add_param(p + ".weight", (1,))
        conv(f"{self.in_name}.conv1", 4 * F, self.cin, 3); prelu(f"{self.in_name}.prelu1")
        conv(f"{self.in_name}.conv2", F, 4 * F, 1); prelu(f"{self.in_name}.prelu2")
        conv("f_block.in_block.conv", F, 2 * F, 1); prelu("f_block.in_block.prelu")
        for g in range(G):                      # nn.ModuleList up_blocks first (drf_net.py:68,79-95)
            if g == 0:
                dconv("f_block.up_blocks.0.deconv", F, F, k); prelu("f_block.up_blocks.0.prelu")
            else:
                conv(f"f_block.up_blocks.{g}.conv1", F, F * (g + 1), 1); prelu(f"f_block.up_blocks.{g}.prelu1")
                dconv(f"f_block.up_blocks.{g}.deconv2", F, F, k); prelu(f"f_block.up_blocks.{g}.prelu2")
        for g in range(G):
            if g == 0:
                conv("f_block.down_blocks.0.conv", F, F, k); prelu("f_block.down_blocks.0.prelu")
            else:
                conv(f"f_block.down_blocks.{g}.conv1", F, F * (g + 1), 1); prelu(f"f_block.down_blocks.{g}.prelu1")
                conv(f"f_block.down_blocks.{g}.conv2", F, F, k); prelu(f"f_block.down_blocks.{g}.prelu2")
        conv("f_block.out_block.conv", F, F * G, 1); prelu("f_block.out_block.prelu")
        if self.variant == "srfb":
            dconv("r_block.deconv1", F, F, k); prelu("r_block.prelu1")
            conv("r_block.conv2", self.cout, F, 3)
            self.out_levels, self.last_name = 0, "r_block.conv2"
        elif self.r == 3:
            conv("out_block.conv1", 9 * F, F, 3); conv("out_block.conv2", self.cout, F, 3)
            self.out_levels, self.last_name = 1, "out_block.conv2"
        else:
            n = {2: 1, 4: 2, 8: 3}[self.r]
            for i in range(n):
                conv(f"out_block.conv{i + 1}", 4 * F, F, 3)
            conv(f"out_block.conv{n + 1}", self.cout, F, 3)
            self.out_levels, self.last_name = n, f"out_block.conv{n + 1}"

    # names of the projection layers of group g
    def up_names(self, g):
        return (None, "f_block.up_blocks.0.deconv", "f_block.up_blocks.0.prelu") if g == 0 else \
            (f"f_block.up_blocks.{g}.conv1", f"f_block.up_blocks.{g}.deconv2", f"f_block.up_blocks.{g}.prelu2")

    def down_names(self, g):
        return (None, "f_block.down_blocks.0.conv", "f_block.down_blocks.0.prelu") if g == 0 else \
            (f"f_block.down_blocks.{g}.conv1", f"f_block.down_blocks.{g}.conv2", f"f_block.down_blocks.{g}.prelu2")

    # ---- helpers ---------------------------------------------------------------------------
    def _W(self, name):
        return self.params[name + ".weight"]

    def _bias_idx(self, name, n=None, perm=None):
        b = self.params[name + ".bias"]
        n = b.shape[0] if n is None else n
        j = np.arange(n)
        return b.offset + (j if perm is None else perm(j))

    def _jk(self, nt):
        return np.arange(nt).reshape(nt, 1), np.arange(self.kc).reshape(1, self.kc)

    # 1x1 convolution over a (virtual) concatenation of `n_src` F-channel maps
    def _conv1x1_cat(self, lname, wname, n_src, src_c=None, slope=None):
        W = self._W(wname)
        cout = W.shape[0]
        src_c = self.F if src_c is None else src_c
        kb = src_c // self.kc
        j, k = self._jk(cout)
        taps, slabs = [], []
        for s in range(n_src):
            for b in range(kb):
                taps.append((s, 0, 0, b * self.kc))
                slabs.append(W.idx(j, s * src_c + b * self.kc + k, 0, 0))
        self.fwd[lname] = Layer(lname, TapTable(self.kc, cout, [(0, taps)]), slabs, cout,
                                self._bias_idx(wname), slope=slope)

    # data-gradient of 1x1 convolutions w.r.t. one F-channel operand of their concat:
    # consumers = [(weight name, input-channel offset of the operand in that layer's concat)]
    def _dgrad1x1(self, lname, consumers, out_c=None):
        out_c = self.F if out_c is None else out_c
        splits = _split_nt(out_c)
        nt = splits[0][1]
        j, k = self._jk(nt)
        groups, slabs = [], []
        for (o0, _) in splits:
            taps = []
            for s, (wname, ci0) in enumerate(consumers):
                W = self._W(wname)
                for b in range(W.shape[0] // self.kc):
                    taps.append((s, 0, 0, b * self.kc))
                    slabs.append(W.idx(b * self.kc + k, ci0 + o0 + j, 0, 0))
            groups.append((o0, taps))
        self.bwd[lname] = Layer(lname, TapTable(self.kc, nt, groups), slabs, out_c)

    # "up" form: LR map (K = channels of the LR map) -> phase-blocked HR map.
    # Used by the transposed convolution forward and by the strided convolution's data-gradient.
    # widx(a, b, ky, kx) gives flat indices with a = LR-side channel (GEMM K), b = HR-side channel.
    def _up_form(self, lname, widx, store, slope=None, bias_name=None):
        F, s, p, k = self.F, self.s, self.p, self.k
        sig = {}
        for h in range(s):
            sig[h] = tuple(d for d in range(-4, 5) if 0 <= h + p - s * d < k)
        groups_by_sig = {}
        for slot, (hy, wx) in enumerate(self.phases):
            groups_by_sig.setdefault((sig[hy], sig[wx]), []).append(slot)
        groups, slabs = [], []
        max_slots = max(1, MAX_NT // F)
        for (sy, sx), slots in sorted(groups_by_sig.items()):
            slots = sorted(slots)
            runs, cur = [], [slots[0]]
            for sl in slots[1:]:
                if sl == cur[-1] + 1 and len(cur) < max_slots:
                    cur.append(sl)
                else:
                    runs.append(cur); cur = [sl]
            runs.append(cur)
            for run in runs:
                groups.append((run, sy, sx))
        # all groups must share nt: use the smallest run length that divides all
        run_len = min(len(g[0]) for g in groups)
        norm = []
        for run, sy, sx in groups:
            for i in range(0, len(run), run_len):
                assert len(run[i:i + run_len]) == run_len
                norm.append((run[i:i + run_len], sy, sx))
        nt = run_len * F
        table_groups = []
        jj = np.arange(nt).reshape(nt, 1)
        kk = np.arange(self.kc).reshape(1, self.kc)
        for run, sy, sx in norm:
            hy = np.array([self.phases[sl][0] for sl in run])[jj // F]
            wx = np.array([self.phases[sl][1] for sl in run])[jj // F]
            taps = []
            for dY in sy:
                for dX in sx:
                    for b in range(self.kb):
                        taps.append((0, dY, dX, b * self.kc))
                        slabs.append(widx(b * self.kc + kk, jj % F, hy + p - s * dY, wx + p - s * dX))
            table_groups.append((run[0] * F, taps))
        bias = None
        if bias_name is not None:
            bias = self._bias_idx(bias_name, len(self.phases) * F, perm=lambda q: q % F)
        store[lname] = Layer(lname, TapTable(self.kc, nt, table_groups), slabs, len(self.phases) * F, bias, slope=slope)

    # "down" form: phase-blocked HR map -> LR map (K runs over taps x HR channels).
    # Used by the strided convolution forward and by the transposed convolution's data-gradient.
    # widx(a, b, ky, kx): a = LR-side (output) channel, b = HR-side channel (GEMM K).
    def _down_form(self, lname, widx, store, slope=None, bias_name=None):
        F, s, p, k = self.F, self.s, self.p, self.k
        j, kk = self._jk(F)
        taps, slabs = [], []
        for ky in range(k):
            dY, hy = divmod(ky - p, s)
            for kx in range(k):
                dX, wx = divmod(kx - p, s)
                slot = self.slot_of[(hy, wx)]
                for b in range(self.kb):
                    taps.append((0, dY, dX, slot * F + b * self.kc))
                    slabs.append(widx(j, b * self.kc + kk, ky, kx))
        bias = self._bias_idx(bias_name) if bias_name is not None else None
        store[lname] = Layer(lname, TapTable(self.kc, F, [(0, taps)]), slabs, F, bias, slope=slope)

    # output-block level: 3x3 convolution F -> m*F on the map of `in_slots` phase slots; the m*F
    # outputs of a pixel are stored phase-major ((i,j),c) = the next level's slots.
    def _out_level(self, level, wname):
        F = self.F
        W = self._W(wname)
        m = W.shape[0] // F                    # 4, or 9 for r == 3
        q = int(round(m ** 0.5))
        in_r = 2 ** level                      # resolution factor of the input map
        in_phases = phase_table(in_r)
        in_slot = {yx: i for i, yx in enumerate(in_phases)}
        # packed output channel jj=(i*q+j)*F + c  <->  torch channel c*m + i*q + j  (PixelShuffle)
        perm = lambda jj: (jj % F) * m + jj // F
        splits = _split_nt(m * F)
        nt = splits[0][1]
        jj, kk = self._jk(nt)
        fgroups, fslabs = [], []
        for slot, (py, px) in enumerate(in_phases):
            for (c_lo, _) in splits:
                taps = []
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        dY, qy = divmod(py + dy, in_r)
                        dX, qx = divmod(px + dx, in_r)
                        for b in range(self.kb):
                            taps.append((0, dY, dX, in_slot[(qy, qx)] * F + b * self.kc))
                            fslabs.append(W.idx(perm(c_lo + jj), b * self.kc + kk, dy + 1, dx + 1))
                fgroups.append((slot * m * F + c_lo, taps))
        n_slots = len(in_phases)
        bias = self._bias_idx(wname, n_slots * m * F, perm=lambda t: perm(t % (m * F)))
        lname = f"out{level + 1}"
        self.fwd[lname] = Layer(lname, TapTable(self.kc, nt, fgroups), fslabs, n_slots * m * F, bias)
        # data-gradient: d_in(P, ci) = sum_{ky,kx,co} dz(P - (ky-1,kx-1), co) W[co,ci,ky,kx]
        kb_out = (m * F) // self.kc
        if in_r >= 2 and n_slots % 4 == 0 and 4 * F <= MAX_NT:
            # The four phase slots of a 2x2 block (nested order: slots 4b .. 4b+3) are produced TOGETHER (nt = 4F): they
            # read the same 4x4 neighbourhood of dz positions, each with its own kernel offset (or none: structural
            # zero rows).  One A tile then feeds 4F output channels instead of F - the N = 64 form ran at a third of the
            # tensor peak (smem operand reads per MMA) and loaded 144 A tiles per pixel tile instead of 64.
            jj, k2 = self._jk(4 * F)
            oi, oj, ci = (jj // F) // 2, (jj // F) % 2, jj % F          # output slot (i, j) of the block, input channel
            bgroups, bslabs = [], []
            parents = phase_table(in_r // 2)
            for b, (Py, Px) in enumerate(parents):
                assert [in_phases[4 * b + 2 * i + j] for i in (0, 1) for j in (0, 1)] == \
                    [(2 * Py + i, 2 * Px + j) for i in (0, 1) for j in (0, 1)]
                taps = []
                rows = sorted(range(2 * Py - 1, 2 * Py + 3), key=lambda R: (R % in_r, R // in_r))
                for C in range(2 * Px - 1, 2 * Px + 3):
                    dX, qx = divmod(C, in_r)
                    for bk in range(kb_out):
                        for R in rows:      # same row phase -> consecutive row shifts, ascending (shared A box)
                            dY, qy = divmod(R, in_r)
                            taps.append((0, dY, dX, in_slot[(qy, qx)] * m * F + bk * self.kc))
                            ky, kx = 2 * Py + oi - R + 1, 2 * Px + oj - C + 1
                            ok = (ky >= 0) & (ky <= 2) & (kx >= 0) & (kx <= 2)
                            idx = W.idx(perm(bk * self.kc + k2), ci, np.clip(ky, 0, 2), np.clip(kx, 0, 2))
                            bslabs.append(np.where(ok, idx, -1))
                bgroups.append((4 * b * F, taps))
            # (9 of the 16 source positions of a block carry a kernel tap for a given output slot)
            self.bwd[lname] = Layer(lname, TapTable(self.kc, 4 * F, bgroups, useful=9.0 / 16.0), bslabs, n_slots * F)
            return
        j2, k2 = self._jk(F)
        bgroups, bslabs = [], []
        for slot, (py, px) in enumerate(in_phases):
            taps = []
            for ky in (2, 1, 0):        # row shifts ascending with the slab index: the tensor-core kernel then shares
                for kx in (2, 1, 0):    # one taller A box between the taps of a column (tapgemm_tc2.cu build_columns)
                    dY, qy = divmod(py - (ky - 1), in_r)
                    dX, qx = divmod(px - (kx - 1), in_r)
                    for b in range(kb_out):
                        taps.append((0, dY, dX, in_slot[(qy, qx)] * m * F + b * self.kc))
                        bslabs.append(W.idx(perm(b * self.kc + k2), j2, ky, kx))
            bgroups.append((slot * F, taps))
        self.bwd[lname] = Layer(lname, TapTable(self.kc, F, bgroups), bslabs, n_slots * F)

    # ---- all layers ------------------------------------------------------------------------
    def _build_layers(self):
        F, G, P = self.F, self.G, self.params
        # in_block.conv1 is the dedicated first-layer kernel (K = 9*Cin).
        self._conv1x1_cat("in2", f"{self.in_name}.conv2", 1, src_c=4 * F, slope=P[f"{self.in_name}.prelu2.weight"])
        self._dgrad1x1("in2", [(f"{self.in_name}.conv2", 0)], out_c=4 * F)   # dz_in -> d(a1)
        self._conv1x1_cat("fin", "f_block.in_block.conv", 2, slope=P["f_block.in_block.prelu.weight"])
        # d(in) from the first concat operand (+ second operand at t == 0 where hidden == in)
        self._dgrad1x1("fin_in", [("f_block.in_block.conv", 0)])
        self._dgrad1x1("fin_in0", [("f_block.in_block.conv", 0), ("f_block.in_block.conv", F)])
        self._dgrad1x1("fin_hid", [("f_block.in_block.conv", F)])
        for g in range(G):
            uc1, udc, upr = self.up_names(g)
            dc1, dsc, dpr = self.down_names(g)
            if g > 0:
                self._conv1x1_cat(f"up{g}_c1", uc1, g + 1, slope=P[f"f_block.up_blocks.{g}.prelu1.weight"])
                self._conv1x1_cat(f"dn{g}_c1", dc1, g + 1, slope=P[f"f_block.down_blocks.{g}.prelu1.weight"])
            WT = self._W(udc)   # ConvTranspose2d weight [Cin, Cout, k, k]
            self._up_form(f"up{g}_dc", lambda a, b, ky, kx, WT=WT: WT.idx(a, b, ky, kx), self.fwd,
                          slope=P[upr + ".weight"], bias_name=udc)
            # dgrad of the transposed conv: LR-side channel = its Cin (output here), HR-side = Cout
            self._down_form(f"up{g}_dc", lambda a, b, ky, kx, WT=WT: WT.idx(a, b, ky, kx), self.bwd)
            Wc = self._W(dsc)   # Conv2d weight [Cout, Cin, k, k]
            self._down_form(f"dn{g}_sc", lambda a, b, ky, kx, Wc=Wc: Wc.idx(a, b, ky, kx), self.fwd,
                            slope=P[dpr + ".weight"], bias_name=dsc)
            # dgrad of the strided conv: LR-side channel = its Cout (GEMM K), HR-side = Cin
            self._up_form(f"dn{g}_sc", lambda a, b, ky, kx, Wc=Wc: Wc.idx(a, b, ky, kx), self.bwd)
        self._conv1x1_cat("fout", "f_block.out_block.conv", G, slope=P["f_block.out_block.prelu.weight"])
        # gathered data-gradients of the LR / HR feature lists (every consumer is a 1x1 conv):
        #   lr_j (j>=1): f_block.out_block.conv slice j-1, up_blocks[g].conv1 slice j for g >= max(j,1)
        #   lr_0       : up_blocks[g].conv1 slice 0 for g >= 1        (+ the deconv of group 0)
        #   hr_j       : down_blocks[g].conv1 slice j for g >= max(j,1) (+ the strided conv of group 0)
        for j in range(G + 1):
            cons = []
            if j >= 1:
                cons.append(("f_block.out_block.conv", (j - 1) * F))
            for g in range(max(j, 1), G):
                cons.append((f"f_block.up_blocks.{g}.conv1", j * F))
            if cons:
                self._dgrad1x1(f"lr{j}", cons)
        for j in range(G):
            cons = [(f"f_block.down_blocks.{g}.conv1", j * F) for g in range(max(j, 1), G)]
            if cons:
                self._dgrad1x1(f"hr{j}", cons)
        if self.variant == "srfb":
            WT = self._W("r_block.deconv1")
            self._up_form("rdc", lambda a, b, ky, kx, WT=WT: WT.idx(a, b, ky, kx), self.fwd,
                          slope=P["r_block.prelu1.weight"], bias_name="r_block.deconv1")
            self._down_form("rdc", lambda a, b, ky, kx, WT=WT: WT.idx(a, b, ky, kx), self.bwd)
        elif self.r == 3:
            self._out_level(0, "out_block.conv1")
        else:
            for lv in range(self.out_levels):
                self._out_level(lv, f"out_block.conv{lv + 1}")

    # ---- precision='bf16x3' (ops.SplitOps): three slabs per tap ------------------------------
    def split_index(self, which):
        """packing map of the strict tensor-core mode for `which` in ('fwd', 'bwd'): every layer's slabs tripled group by
        group as [wh of the group's taps | wh again | wl] (bit 30 of an index = the low-order bf16 part of the parameter:
        vsr_gather_split), matching SplitOps.table3; layer L's slabs start at 3 * L.w_off"""
        if which not in self._split_idx:
            assert self.bf16 and self.kc == 64, "bf16x3 mode uses the bf16 slab layout"
            base = self.fwd_w_idx if which == "fwd" else self.bwd_w_idx
            store = self.fwd if which == "fwd" else self.bwd
            chunks = []
            for L in store.values():
                slab = L.table.nt * 64
                begin = L.w_off
                for _, taps in L.table.groups:
                    g = base[begin:begin + len(taps) * slab].astype(np.int64)
                    chunks += [g, g, np.where(g >= 0, g | (1 << 30), -1)]
                    begin += len(taps) * slab
                assert begin == L.w_off + L.w_numel
            self._split_idx[which] = np.concatenate(chunks).astype(np.int32)
        return self._split_idx[which]

    # ---- gradient buckets ------------------------------------------------------------------
    def grad_buckets(self, n_buckets=3):
        """Contiguous ranges [lo, hi) of the flat gradient bucket with the tap-GEMM layers whose weight and bias
        gradients land inside each: the engine finishes the ranges one after the other at the end of backward, so
        that the data-parallel trainer can all-reduce range k while the weight gradients of range k+1 are still
        being computed.  The ranges tile [0, n_params); a layer's parameters never straddle a boundary."""
        spans = []
        for L in self.fwd.values():
            idx = np.concatenate([np.asarray(sl).reshape(-1) for sl in L.slabs])
            idx = idx[idx >= 0]
            lo, hi = int(idx.min()), int(idx.max()) + 1
            if L.bias_idx is not None:
                b = np.asarray(L.bias_idx)
                b = b[b >= 0]
                lo, hi = min(lo, int(b.min())), max(hi, int(b.max()) + 1)
            spans.append((lo, hi, L.name))
        spans.sort()
        for (l0, h0, n0), (l1, h1, n1) in zip(spans, spans[1:]):
            assert h0 <= l1, f"layers {n0} and {n1} share parameters: gradient ranges would overlap"
        total = sum(h - l for l, h, _ in spans)
        buckets, cur, acc, lo = [], [], 0, 0
        for i, (l, h, name) in enumerate(spans):
            cur.append(name)
            acc += h - l
            last = i == len(spans) - 1
            if last or (len(buckets) < n_buckets - 1 and acc >= total * (len(buckets) + 1) / n_buckets):
                hi = self.n_params if last else spans[i + 1][0]
                buckets.append((lo, hi, cur))
                cur, lo = [], hi
        return buckets

    # ---- packing maps ----------------------------------------------------------------------
    def _finalize(self):
        def pack(store):
            w_chunks, b_chunks, w_off, b_off = [], [], 0, 0
            for L in store.values():
                L.w_off, L.b_off = w_off, b_off
                arr = np.stack(L.slabs).astype(np.int64)                 # [T, nt, kc]
                if self.bf16:
                    assert self.kc == 64
                    nt = L.table.nt
                    jj = np.arange(nt).reshape(nt, 1)
                    kk = np.arange(64).reshape(1, 64)
                    pos = (jj * 64 + (((kk >> 3) ^ (jj & 7)) << 3) + (kk & 7)).reshape(-1)
                    sw = np.empty((arr.shape[0], nt * 64), dtype=np.int64)
                    sw[:, pos] = arr.reshape(arr.shape[0], -1)
                    arr = sw
                w_chunks.append(arr.reshape(-1))
                w_off += L.w_numel
                if L.bias_idx is not None:
                    b_chunks.append(np.asarray(L.bias_idx, dtype=np.int64))
                    b_off += len(L.bias_idx)
                else:
                    L.b_off = -1
            cat = lambda ch: np.concatenate(ch) if ch else np.zeros(0, np.int64)
            return cat(w_chunks).astype(np.int32), cat(b_chunks).astype(np.int32)

        self.fwd_w_idx, self.fwd_b_idx = pack(self.fwd)
        self.bwd_w_idx, _ = pack(self.bwd)
        self._split_idx = {}
        # weight-gradient un-packing: packed dW is in PLAIN [T][nt][kc] order at the same offsets.
        # For every parameter element list the packed positions that accumulate into it.
        pidx_all, ppos_all = [], []
        for L in self.fwd.values():
            arr = np.stack(L.slabs).astype(np.int64).reshape(-1)
            where = np.nonzero(arr >= 0)[0]
            pidx_all.append(arr[where])
            ppos_all.append(where + L.w_off)
        pidx_all, ppos_all = np.concatenate(pidx_all), np.concatenate(ppos_all)
        order = np.argsort(pidx_all, kind="stable")
        pk, qk = pidx_all[order], ppos_all[order]
        first = np.r_[True, pk[1:] != pk[:-1]]
        start = np.maximum.accumulate(np.where(first, np.arange(len(pk)), 0))
        rank = np.arange(len(pk)) - start
        self.unpack_passes = []
        for d in range(int(rank.max()) + 1):
            sel = rank == d
            idx = np.full(self.n_params, -1, dtype=np.int32)
            idx[pk[sel]] = qk[sel]
            lo, hi = int(pk[sel].min()), int(pk[sel].max()) + 1
            self.unpack_passes.append((lo, idx[lo:hi].copy()))
        # bias-gradient un-packing: column sums land in the first period of each packed bias
        bidx = np.full(self.n_params, -1, dtype=np.int32)
        for L in self.fwd.values():
            if L.bias_idx is not None:
                bi = np.asarray(L.bias_idx, dtype=np.int64)
                valid = bi >= 0
                if not valid.all():      # padded / windowed output channels (-1): the period is the whole vector
                    L.bias_c = len(bi)
                    bidx[bi[valid]] = L.b_off + np.nonzero(valid)[0]
                    continue
                L.bias_c = int(len(np.unique(L.bias_idx)))
                assert np.array_equal(L.bias_idx[:L.bias_c], L.bias_idx[L.bias_c:2 * L.bias_c]) or L.bias_c == L.out_c
                bidx[np.asarray(L.bias_idx[:L.bias_c], dtype=np.int64)] = L.b_off + np.arange(L.bias_c)
        self.bias_unpack_idx = bidx
        self.fwd_w_numel = int(sum(L.w_numel for L in self.fwd.values()))
        self.bwd_w_numel = int(sum(L.w_numel for L in self.bwd.values()))
        self.fwd_b_numel = int(len(self.fwd_b_idx))
