"""Thin host runtime over the C ABI: device buffers and streams come from torch (plumbing only),
every computation is a call into libapax_b200.so.  No oracle, no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import DISP_DRVEC, DISP_GAS, DISP_MINIMAGE, DISP_OFFSETS, GmConfig, check, lib

MODES = {"gas": DISP_GAS, "offsets": DISP_OFFSETS, "minimage": DISP_MINIMAGE, "drvec": DISP_DRVEC}
N_FEATURES = 360
FEATURE_COUNTS = (5, 20, 35, 50, 85, 160, 235, 360)   # c0 | c0..c1 | ... | c0..c7 (gaussian_moment_descriptor.py:56-101)


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "runtime", "no CUDA device: apax_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def to_device(x, dtype: torch.dtype) -> torch.Tensor:
    """Host array or tensor -> contiguous CUDA tensor of `dtype` (no copy if already there)."""
    dev = _require_cuda()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(dev)


def default_config(n_out: int = 1, **overrides) -> dict:
    """Flax-module defaults of the GMNN path (SURVEY.md §8 header; apax/config/model_config.py:147-227)."""
    cfg = dict(n_species=119, n_radial=5, n_basis=7, n_contr=8, nn=[256, 256], n_out=n_out, use_ntk=False,
               use_embed_norm=True, mask_atoms=True, r_min=0.5, r_max=6.0, basis_kind=0)
    cfg.update(overrides)
    return cfg


class DeviceParams:
    """Device-resident copy of an apax parameter tree (SURVEY.md §5 layout), owned by the C library.

    Accepts {"params": {"energy_model": {...}}} or the inner "energy_model" dict; arrays may be
    numpy / torch / anything np.asarray understands.
    """

    def __init__(self, params: dict, config: Optional[dict] = None):
        _require_cuda()
        tree = params.get("params", params)
        tree = tree.get("energy_model", tree)
        emb = np.ascontiguousarray(np.asarray(tree["representation"]["radial_fn"]["atomic_type_embedding"]), dtype=np.float32)
        ro = tree["readout"]
        layers = [ro[f"dense_{i}"] for i in range(len(ro))]
        if len(layers) != 3:
            raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "DeviceParams", "readout must be [256, 256] + output layer")
        w = [np.ascontiguousarray(np.asarray(l["w"]), dtype=np.float32) for l in layers]
        b = [np.ascontiguousarray(np.asarray(l["b"]), dtype=np.float32) for l in layers]
        n_out = int(w[2].shape[1])
        cfg = default_config(n_out=n_out)
        if config:
            cfg.update(config)
        cfg["n_out"] = n_out
        cfg["n_species"] = int(emb.shape[0])
        if emb.shape != (cfg["n_species"], cfg["n_species"], cfg["n_radial"], cfg["n_basis"]):
            raise _lib.ApaxB200Error(_lib.ERR_INVALID_ARG, "DeviceParams", f"embedding shape {emb.shape}")
        if not 1 <= int(cfg["n_contr"]) <= 8:
            raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "DeviceParams", f"n_contr = {cfg['n_contr']}")
        self.n_features = FEATURE_COUNTS[int(cfg["n_contr"]) - 1]
        if w[0].shape != (self.n_features, 256) or w[1].shape != (256, 256) or w[2].shape[0] != 256:
            raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "DeviceParams",
                                     f"dense shapes must be {self.n_features}x256, 256x256, 256xn_out (got {w[0].shape}, {w[1].shape}, {w[2].shape})")
        ss = tree["scale_shift"]
        S = cfg["n_species"]
        scale = np.ascontiguousarray(np.broadcast_to(np.asarray(ss["scale_per_element"], dtype=np.float64).reshape(-1), (S,)))
        shift = np.ascontiguousarray(np.broadcast_to(np.asarray(ss["shift_per_element"], dtype=np.float64).reshape(-1), (S,)))
        self.config = cfg
        self.n_out = n_out
        c = GmConfig(n_species=S, n_radial=cfg["n_radial"], n_basis=cfg["n_basis"], n_contr=cfg["n_contr"],
                     n_hidden1=cfg["nn"][0], n_hidden2=cfg["nn"][1], n_out=n_out, use_ntk=int(cfg["use_ntk"]),
                     use_embed_norm=int(cfg["use_embed_norm"]), mask_atoms=int(cfg["mask_atoms"]),
                     r_min=cfg["r_min"], r_max=cfg["r_max"], basis_kind=int(cfg.get("basis_kind", 0)))
        self._keep = (emb, w, b, scale, shift)
        handle = C.c_void_p()
        # pairwise empirical corrections (EnergyModel.corrections, apax/nn/models.py:127-130): flax names the list entries
        # corrections_0, corrections_1, ...; a ZBL block holds "coefficients", an exponential block "rep_prefactor"
        corr, corr_keep = None, []
        for key in sorted(k for k in tree if str(k).startswith("corrections_")):
            blk = tree[key]
            corr = corr or _lib.GmCorrections()
            if "coefficients" in blk:
                if cfg.get("zbl_r_max") is None:
                    raise _lib.ApaxB200Error(_lib.ERR_INVALID_ARG, "DeviceParams", "ZBL parameters without zbl_r_max in the config")
                co = np.asarray(blk["coefficients"], dtype=np.float32).reshape(4)
                corr.zbl_on, corr.zbl_r_max = 1, float(cfg["zbl_r_max"])
                corr.zbl_rep_scale = float(np.asarray(blk["rep_scale"], dtype=np.float32).reshape(-1)[0])
                for q in range(4):
                    corr.zbl_coefficients[q] = float(co[q])
            elif "rep_prefactor" in blk:
                if cfg.get("exp_r_max") is None:
                    raise _lib.ApaxB200Error(_lib.ERR_INVALID_ARG, "DeviceParams", "repulsion parameters without exp_r_max in the config")
                rs = np.ascontiguousarray(np.asarray(blk["rep_scale"], dtype=np.float32).reshape(-1))
                pf = np.ascontiguousarray(np.asarray(blk["rep_prefactor"], dtype=np.float32).reshape(-1))
                if rs.shape[0] != S or pf.shape[0] != S:
                    raise _lib.ApaxB200Error(_lib.ERR_INVALID_ARG, "DeviceParams", "rep_scale / rep_prefactor must have n_species entries")
                corr_keep += [rs, pf]
                corr.exp_on, corr.exp_r_max = 1, float(cfg["exp_r_max"])
                corr.exp_rep_scale_host, corr.exp_rep_prefactor_host = rs.ctypes.data, pf.ctypes.data
            else:
                raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "DeviceParams", f"empirical correction {key} (only ZBL / exponential)")

        def hp(a):
            return a.ctypes.data_as(C.c_void_p)

        check(lib().apax_gm_params_create(C.byref(c), hp(emb), hp(w[0]), hp(b[0]), hp(w[1]), hp(b[1]), hp(w[2]), hp(b[2]),
                                          hp(scale), hp(shift), C.byref(handle)), "apax_gm_params_create")
        self.handle = handle
        # implementation switches live in the handle (apax_gm_params_set_option), never in the environment
        if cfg.get("readout") is not None:
            check(lib().apax_gm_params_set_option(handle, _lib.OPT_READOUT, _lib.READOUT_KINDS[str(cfg["readout"])]),
                  "apax_gm_params_set_option(readout)")
        self.n_force_heads = n_out            # rows of the `forces` output: one per head, or one with mean_forces
        if cfg.get("mean_forces"):
            check(lib().apax_gm_params_set_option(handle, _lib.OPT_MEAN_FORCES, 1), "apax_gm_params_set_option(mean_forces)")
            self.n_force_heads = 1
        if cfg.get("contract_split") is not None:
            check(lib().apax_gm_params_set_option(handle, _lib.OPT_CONTRACT_SPLIT, int(cfg["contract_split"])),
                  "apax_gm_params_set_option(contract_split)")
        if cfg.get("pair_block") is not None:
            check(lib().apax_gm_params_set_option(handle, _lib.OPT_PAIR_BLOCK, int(cfg["pair_block"])),
                  "apax_gm_params_set_option(pair_block)")
        if corr is not None:
            check(lib().apax_gm_params_set_corrections(handle, C.byref(corr)), "apax_gm_params_set_corrections")

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                lib().apax_gm_params_destroy(h)
            except Exception:
                pass
            self.handle = None


class Workspace:
    """Caller-owned scratch for the gm_* entry points, grown on demand and reused."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=_require_cuda())
        return self.buf

    @staticmethod
    def aligned_ptr(buf: torch.Tensor) -> int:
        return (buf.data_ptr() + 255) & ~255


_default_ws = Workspace()
_nl_ws = Workspace()


def _gm_ws(n_atoms: int, n_pairs: int, n_out: int, ws: Optional[Workspace]):
    ws = ws or _default_ws
    nbytes = lib().apax_gm_workspace_bytes(n_atoms, n_pairs, n_out)
    buf = ws.get(nbytes)
    return C.c_void_p(Workspace.aligned_ptr(buf)), C.c_size_t(nbytes)


def energy_forces(params: DeviceParams, R, Z, idx, box, offsets, mode: str = "offsets", forces: bool = True,
                  workspace: Optional[Workspace] = None, stress: bool = False):
    """EnergyDerivativeModel.apply (apax/nn/models.py:146-171) on the GPU.

    Returns (energy f64 tensor [n_out], forces f64 tensor [n_out, N, 3] or None), asynchronous on the
    current torch stream.
    """
    dev = _require_cuda()
    R = to_device(R, torch.float64)
    Z = to_device(Z, torch.int32)
    idx = to_device(idx, torch.int32)
    n_atoms, n_pairs = int(R.shape[0]), int(idx.shape[1])
    box_d = to_device(np.zeros((3, 3)) if box is None else box, torch.float64)
    if box_d.numel() == 3:  # apax passes np.array([0,0,0]) for gas phase
        box_d = torch.diag(box_d)
    off_d = None if offsets is None else to_device(offsets, torch.float64)
    if off_d is not None and off_d.dim() == 1:  # simulate.py:71 passes a broadcast [0,0,0]
        off_d = None if not bool(off_d.any()) else off_d.expand(n_pairs, 3).contiguous()
    energy = torch.empty(params.n_out, dtype=torch.float64, device=dev)
    F = torch.empty((params.n_force_heads, n_atoms, 3), dtype=torch.float64, device=dev) if forces else None
    wp, wb = _gm_ws(n_atoms, n_pairs, params.n_out, workspace)
    check(lib().apax_gm_energy_forces(_stream_ptr(), params.handle, _ptr(R), _ptr(Z), _ptr(idx), _ptr(box_d), _ptr(off_d),
                                      n_atoms, n_pairs, MODES[mode], _ptr(energy), _ptr(F), wp, wb),
          "apax_gm_energy_forces")
    if stress:
        if mode not in ("minimage", "offsets") or not forces:
            raise _lib.ApaxB200Error(_lib.ERR_UNSUPPORTED, "energy_forces", "stress needs a periodic displacement (with forces)")
        st = torch.empty((3, 3), dtype=torch.float64, device=dev)
        if mode == "minimage":
            check(lib().apax_gm_stress_times_vol(_stream_ptr(), params.handle, n_atoms, n_pairs, _ptr(st), wp, wb),
                  "apax_gm_stress_times_vol")
        else:
            # the reference's disp_fn under a perturbation (apax/layers/distances.py:12-22): dU/d(eps) at doubled displacements
            box2 = box_d * 2.0
            e2, F2 = torch.empty_like(energy), torch.empty_like(F)
            check(lib().apax_gm_energy_forces(_stream_ptr(), params.handle, _ptr(R), _ptr(Z), _ptr(idx), _ptr(box2), _ptr(off_d),
                                              n_atoms, n_pairs, MODES[mode], _ptr(e2), _ptr(F2), wp, wb),
                  "apax_gm_energy_forces")
            check(lib().apax_gm_stress_times_vol_offsets(_stream_ptr(), params.handle, _ptr(R), _ptr(box_d), n_atoms, n_pairs,
                                                         _ptr(st), wp, wb), "apax_gm_stress_times_vol_offsets")
        return energy, F, st
    return energy, F


class PreparedSystem:
    """A neighbour list prepared once (apax_gm_prepare) and evaluated many times (apax_gm_compute)."""

    def __init__(self, params: DeviceParams, Z, idx, n_atoms: Optional[int] = None, n_central: Optional[int] = None,
                 workspace: Optional[Workspace] = None, forces: Optional[torch.Tensor] = None):
        self.params = params
        self.n_central = n_central
        self.Z = to_device(Z, torch.int32)
        self.idx = to_device(idx, torch.int32)
        self.n_atoms = int(self.Z.shape[0]) if n_atoms is None else n_atoms
        self.n_pairs = int(self.idx.shape[1])
        self.ws = workspace or Workspace()
        self.wp, self.wb = _gm_ws(self.n_atoms, self.n_pairs, params.n_out, self.ws)
        check(lib().apax_gm_prepare(_stream_ptr(), params.handle, _ptr(self.Z), _ptr(self.idx), self.n_atoms, self.n_pairs,
                                    self.wp, self.wb), "apax_gm_prepare")
        dev = self.Z.device
        self.energy = torch.empty(params.n_out, dtype=torch.float64, device=dev)
        # `forces` may be caller-provided memory (e.g. a peer-mapped block for the NVLink halo exchange)
        self.forces = forces if forces is not None else torch.empty((params.n_force_heads, self.n_atoms, 3), dtype=torch.float64, device=dev)
        assert tuple(self.forces.shape) == (params.n_force_heads, self.n_atoms, 3) and self.forces.dtype == torch.float64

    def compute(self, R: torch.Tensor, box: Optional[torch.Tensor], offsets: Optional[torch.Tensor], mode: str,
                forces: bool = True):
        if self.n_central is not None:
            check(lib().apax_gm_compute_partial(_stream_ptr(), self.params.handle, _ptr(R), _ptr(self.Z), _ptr(box),
                                                _ptr(offsets), self.n_atoms, self.n_pairs, self.n_central, MODES[mode],
                                                _ptr(self.energy), _ptr(self.forces if forces else None), self.wp,
                                                self.wb), "apax_gm_compute_partial")
        else:
            check(lib().apax_gm_compute(_stream_ptr(), self.params.handle, _ptr(R), _ptr(self.Z), _ptr(box), _ptr(offsets),
                                        self.n_atoms, self.n_pairs, MODES[mode], _ptr(self.energy),
                                        _ptr(self.forces if forces else None), self.wp, self.wb), "apax_gm_compute")
        return self.energy, (self.forces if forces else None)

    def atomic_energies(self) -> torch.Tensor:
        """Per-atom energies E_i [n_out, n_atoms] f64 of the last compute() (models.py:109-122, before fp64_sum)."""
        out = torch.empty((self.params.n_out, self.n_atoms), dtype=torch.float64, device=self.Z.device)
        check(lib().apax_gm_atomic_energies(_stream_ptr(), self.params.handle, self.n_atoms, self.n_pairs, _ptr(out), self.wp,
                                            self.wb), "apax_gm_atomic_energies")
        return out


class PreparedMinimageSystem(PreparedSystem):
    """Min-image neighbour list built on the device and prepared in the same call (apax_gm_prepare_minimage): positions stay
    on the GPU, the list never takes COO form.  `build(frac, box, cutoff)` may be called again whenever the list must be
    rebuilt; `check()` reads the overflow flag (one small synchronising copy)."""

    def __init__(self, params: DeviceParams, Z, capacity: int, n_central: Optional[int] = None,
                 forces: Optional[torch.Tensor] = None, workspaces: Optional[tuple] = None):
        dev = _require_cuda()
        self.params = params
        self.Z = to_device(Z, torch.int32)
        self.n_atoms = int(self.Z.shape[0])
        self.n_central = n_central
        self.n_pairs = int(capacity)          # the workspace is carved for the capacity; compute calls pass the same number
        self.ws, self.nl_ws = workspaces if workspaces is not None else (Workspace(), Workspace())   # reusable across re-creations
        self.wp, self.wb = _gm_ws(self.n_atoms, self.n_pairs, params.n_out, self.ws)
        nbytes = lib().apax_nl_workspace_bytes(self.n_atoms, self.n_pairs)
        self.nlp, self.nlb = C.c_void_p(Workspace.aligned_ptr(self.nl_ws.get(nbytes))), C.c_size_t(nbytes)
        self.n_found = torch.zeros(1, dtype=torch.int32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.uint8, device=dev)
        self.energy = torch.empty(params.n_out, dtype=torch.float64, device=dev)
        self.forces = forces if forces is not None else torch.empty((params.n_force_heads, self.n_atoms, 3), dtype=torch.float64, device=dev)
        assert tuple(self.forces.shape) == (params.n_force_heads, self.n_atoms, 3) and self.forces.dtype == torch.float64

    def build(self, frac: torch.Tensor, box: torch.Tensor, cutoff: float) -> None:
        check(lib().apax_gm_prepare_minimage(_stream_ptr(), self.params.handle, _ptr(frac), _ptr(self.Z), _ptr(box), self.n_atoms,
                                             self.n_atoms if self.n_central is None else self.n_central, float(cutoff), self.n_pairs,
                                             _ptr(self.n_found), _ptr(self.overflow), self.wp, self.wb, self.nlp, self.nlb),
              "apax_gm_prepare_minimage")

    def check(self) -> int:
        """Entries found; raises if the list did not fit the capacity."""
        nf, ov = int(self.n_found.item()), int(self.overflow.item())
        if ov or nf > self.n_pairs:
            raise _lib.ApaxB200Error(_lib.ERR_OVERFLOW, "PreparedMinimageSystem", f"{nf} list entries exceed the capacity {self.n_pairs}")
        return nf


def descriptor(params: DeviceParams, R, Z, idx, box=None, offsets=None, mode: str = "gas",
               workspace: Optional[Workspace] = None) -> torch.Tensor:
    """GaussianMomentDescriptor features [N, 360] f32; with mode="drvec", R is dr_vec [P,3]."""
    dev = _require_cuda()
    R = to_device(R, torch.float64)
    Z = to_device(Z, torch.int32)
    idx = to_device(idx, torch.int32)
    n_atoms, n_pairs = int(Z.shape[0]), int(idx.shape[1])
    box_d = None if box is None else to_device(box, torch.float64)
    off_d = None if offsets is None else to_device(offsets, torch.float64)
    g = torch.empty((n_atoms, N_FEATURES), dtype=torch.float32, device=dev)
    wp, wb = _gm_ws(n_atoms, n_pairs, params.n_out, workspace)
    check(lib().apax_gm_descriptor(_stream_ptr(), params.handle, _ptr(R), _ptr(Z), _ptr(idx), _ptr(box_d), _ptr(off_d),
                                   n_atoms, n_pairs, MODES[mode], _ptr(g), wp, wb), "apax_gm_descriptor")
    return g if params.n_features == N_FEATURES else g[:, : params.n_features].contiguous()   # n_contr truncation (:101)


def descriptor_vjp(params: DeviceParams, R, Z, idx, g_bar, box=None, offsets=None, mode: str = "gas", want_dr_vec: bool = True,
                   want_R: bool = False, workspace: Optional[Workspace] = None):
    """VJP of the descriptor (apax_gm_descriptor_vjp): g_bar [N, n_feat] f32 -> (dr_vec_bar [P,3] f64 or None,
    R_bar [N,3] f64 or None).  With mode="drvec", R is dr_vec [P,3] and only dr_vec_bar exists."""
    dev = _require_cuda()
    R = to_device(R, torch.float64)
    Z = to_device(Z, torch.int32)
    idx = to_device(idx, torch.int32)
    n_atoms, n_pairs = int(Z.shape[0]), int(idx.shape[1])
    gb = to_device(g_bar, torch.float32)
    if gb.shape != (n_atoms, params.n_features):
        raise _lib.ApaxB200Error(_lib.ERR_INVALID_ARG, "descriptor_vjp", f"g_bar shape {tuple(gb.shape)}")
    if params.n_features != N_FEATURES:   # n_contr truncation: the dropped features carry a zero cotangent
        full = torch.zeros((n_atoms, N_FEATURES), dtype=torch.float32, device=dev)
        full[:, : params.n_features] = gb
        gb = full
    box_d = None if box is None else to_device(box, torch.float64)
    off_d = None if offsets is None else to_device(offsets, torch.float64)
    d_dr = torch.empty((n_pairs, 3), dtype=torch.float64, device=dev) if want_dr_vec else None
    d_R = torch.empty((n_atoms, 3), dtype=torch.float64, device=dev) if want_R else None
    wp, wb = _gm_ws(n_atoms, n_pairs, params.n_out, workspace)
    check(lib().apax_gm_descriptor_vjp(_stream_ptr(), params.handle, _ptr(R), _ptr(Z), _ptr(idx), _ptr(box_d), _ptr(off_d), n_atoms,
                                       n_pairs, MODES[mode], _ptr(gb), _ptr(d_dr), _ptr(d_R), wp, wb), "apax_gm_descriptor_vjp")
    return d_dr, d_R


# ------------------------------------------------------------------------------------------- lists
def nl_build_minimage(frac, box, cutoff: float, capacity: int):
    """JaxMD Sparse min-image list (partition.py:478-787): idx i32 [2,capacity] padded with N,
    n_found i32 [1], overflow u8 [1] -- all device tensors."""
    dev = _require_cuda()
    frac = to_device(frac, torch.float64)
    box = to_device(box, torch.float64)
    n = int(frac.shape[0])
    idx = torch.empty((2, capacity), dtype=torch.int32, device=dev)
    n_found = torch.zeros(1, dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.uint8, device=dev)
    nbytes = lib().apax_nl_workspace_bytes(n, capacity)
    buf = _nl_ws.get(nbytes)
    check(lib().apax_nl_build_minimage(_stream_ptr(), _ptr(frac), _ptr(box), n, float(cutoff), capacity, _ptr(idx),
                                       _ptr(n_found), _ptr(overflow), C.c_void_p(Workspace.aligned_ptr(buf)),
                                       C.c_size_t(nbytes)), "apax_nl_build_minimage")
    return idx, n_found, overflow


def nl_build_images(positions, cell, cutoff: float, capacity: int, pbc: bool = True, want_offsets: bool = True):
    """vesin-style full list with shifts (ase_calc.py:337-355): idx [2,capacity] zero padded,
    shifts i32 [capacity,3], offsets f64 [capacity,3] = S @ cell, n_found, overflow."""
    dev = _require_cuda()
    pos = to_device(positions, torch.float64)
    cell = to_device(cell, torch.float64)
    n = int(pos.shape[0])
    idx = torch.empty((2, capacity), dtype=torch.int32, device=dev)
    shifts = torch.empty((capacity, 3), dtype=torch.int32, device=dev)
    offsets = torch.empty((capacity, 3), dtype=torch.float64, device=dev) if want_offsets else None
    n_found = torch.zeros(1, dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.uint8, device=dev)
    nbytes = lib().apax_nl_workspace_bytes(n, capacity)
    buf = _nl_ws.get(nbytes)
    check(lib().apax_nl_build_images(_stream_ptr(), _ptr(pos), _ptr(cell), int(bool(pbc)), n, float(cutoff), capacity,
                                     _ptr(idx), _ptr(shifts), _ptr(offsets), _ptr(n_found), _ptr(overflow),
                                     C.c_void_p(Workspace.aligned_ptr(buf)), C.c_size_t(nbytes)), "apax_nl_build_images")
    return idx, shifts, offsets, n_found, overflow


def nl_build_images_batched(positions, cells, struct_ptr, cutoff: float, capacity: int):
    """vesin-style lists for a batch of independent periodic cells (atoms concatenated, global indices)."""
    dev = _require_cuda()
    pos = to_device(positions, torch.float64)
    cells = to_device(cells, torch.float64)
    sp = to_device(struct_ptr, torch.int32)
    n, b = int(pos.shape[0]), int(sp.shape[0]) - 1
    idx = torch.empty((2, capacity), dtype=torch.int32, device=dev)
    shifts = torch.empty((capacity, 3), dtype=torch.int32, device=dev)
    offsets = torch.empty((capacity, 3), dtype=torch.float64, device=dev)
    n_found = torch.zeros(1, dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.uint8, device=dev)
    nbytes = lib().apax_nl_batched_workspace_bytes(n, capacity, b)
    buf = _nl_ws.get(nbytes)
    check(lib().apax_nl_build_images_batched(_stream_ptr(), _ptr(pos), _ptr(cells), _ptr(sp), b, n, float(cutoff), capacity,
                                             _ptr(idx), _ptr(shifts), _ptr(offsets), _ptr(n_found), _ptr(overflow),
                                             C.c_void_p(Workspace.aligned_ptr(buf)), C.c_size_t(nbytes)),
          "apax_nl_build_images_batched")
    return idx, shifts, offsets, n_found, overflow


def energy_forces_batched(params: DeviceParams, R, Z, idx, offsets, struct_ptr, forces: bool = True,
                          workspace: Optional[Workspace] = None):
    """Many independent structures in one call: energy [B, n_out], forces [n_out, N_total, 3]."""
    dev = _require_cuda()
    R = to_device(R, torch.float64)
    Z = to_device(Z, torch.int32)
    idx = to_device(idx, torch.int32)
    sp = to_device(struct_ptr, torch.int32)
    off_d = None if offsets is None else to_device(offsets, torch.float64)
    n_atoms, n_pairs, b = int(R.shape[0]), int(idx.shape[1]), int(sp.shape[0]) - 1
    energy = torch.empty((b, params.n_out), dtype=torch.float64, device=dev)
    F = torch.empty((params.n_force_heads, n_atoms, 3), dtype=torch.float64, device=dev) if forces else None
    wp, wb = _gm_ws(n_atoms, n_pairs, params.n_out, workspace)
    check(lib().apax_gm_energy_forces_batched(_stream_ptr(), params.handle, _ptr(R), _ptr(Z), _ptr(idx), _ptr(off_d), _ptr(sp),
                                              b, n_atoms, n_pairs, _ptr(energy), _ptr(F), wp, wb),
          "apax_gm_energy_forces_batched")
    return energy, F


def ensemble_stats(energy_ens: torch.Tensor, forces_ens: torch.Tensor, want_ensemble: bool = True):
    """ShallowEnsembleModel statistics (apax/nn/models.py:221-263) on the device: energy_ens [B, n_ens], forces_ens
    [n_ens, N, 3] -> (e_mean [B], e_unc [B], f_mean [N,3], f_unc [N,3], forces_ensemble [N,3,n_ens] or None)."""
    dev = _require_cuda()
    b, n_ens = int(energy_ens.shape[0]), int(energy_ens.shape[1])
    n = int(forces_ens.shape[1])
    e_mean = torch.empty(b, dtype=torch.float64, device=dev)
    e_unc = torch.empty(b, dtype=torch.float64, device=dev)
    f_mean = torch.empty((n, 3), dtype=torch.float64, device=dev)
    f_unc = torch.empty((n, 3), dtype=torch.float64, device=dev)
    f_t = torch.empty((n, 3, n_ens), dtype=torch.float64, device=dev) if want_ensemble else None
    check(lib().apax_ensemble_stats(_stream_ptr(), _ptr(energy_ens.contiguous()), _ptr(forces_ens.contiguous()), b, n, n_ens,
                                    _ptr(e_mean), _ptr(e_unc), _ptr(f_mean), _ptr(f_unc), _ptr(f_t)), "apax_ensemble_stats")
    return e_mean, e_unc, f_mean, f_unc, f_t


def nl_needs_rebuild(frac, ref, box, dr_threshold: float) -> torch.Tensor:
    dev = _require_cuda()
    frac = to_device(frac, torch.float64)
    ref = to_device(ref, torch.float64)
    box = to_device(box, torch.float64)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().apax_nl_needs_rebuild(_stream_ptr(), _ptr(frac), _ptr(ref), _ptr(box), int(frac.shape[0]),
                                      float(dr_threshold), _ptr(flag)), "apax_nl_needs_rebuild")
    return flag


class HostContext:
    """apax_ctx_*: host buffers in, host buffers out (the reference's ASECalculator.calculate flow)."""

    def __init__(self, params: DeviceParams, max_atoms: int, max_pairs: int):
        _require_cuda()
        self.params = params
        self.handle = C.c_void_p()
        check(lib().apax_ctx_create(params.handle, max_atoms, max_pairs, C.byref(self.handle)), "apax_ctx_create")

    def calculate(self, positions: np.ndarray, numbers: np.ndarray, cell: np.ndarray, dr_threshold: float = 0.0,
                  rebuild: bool = False, forces: bool = True, out_forces: Optional[np.ndarray] = None):
        """positions / numbers / out_forces may be views of page-locked memory (e.g. `torch.empty(...).pin_memory().numpy()`),
        in which case the library copies from / into them directly without its staging buffer."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        num = np.ascontiguousarray(numbers, dtype=np.int32)
        cel = np.ascontiguousarray(cell, dtype=np.float64)
        n = pos.shape[0]
        e = np.empty(self.params.n_out, dtype=np.float64)
        if out_forces is not None:
            assert out_forces.dtype == np.float64 and out_forces.shape == (self.params.n_force_heads, n, 3) and out_forces.flags.c_contiguous
        f = (out_forces if out_forces is not None else np.empty((self.params.n_force_heads, n, 3), dtype=np.float64)) if forces else None
        npairs = C.c_int64(0)
        code = lib().apax_ctx_calculate(self.handle, pos.ctypes.data_as(C.c_void_p), num.ctypes.data_as(C.c_void_p),
                                        cel.ctypes.data_as(C.c_void_p), n, float(dr_threshold), int(rebuild),
                                        e.ctypes.data_as(C.c_void_p),
                                        f.ctypes.data_as(C.c_void_p) if forces else C.c_void_p(0), C.byref(npairs))
        self.n_pairs = npairs.value
        check(code, "apax_ctx_calculate")
        return e, f

    def stress(self) -> np.ndarray:
        """`stress` of the structure evaluated by the last calculate(forces=True): ProcessStress applied to
        stress_times_vol (apax/md/function_transformations.py:140-159), [3,3] f64."""
        st = np.empty((3, 3), dtype=np.float64)
        check(lib().apax_ctx_stress(self.handle, st.ctypes.data_as(C.c_void_p)), "apax_ctx_stress")
        return st

    def close(self):
        if self.handle is not None and self.handle.value:
            lib().apax_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaView:
    """__cuda_array_interface__ carrier: lets torch view device memory this library allocated (apax_peer_alloc)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_view(ptr: int, shape, dtype: torch.dtype) -> torch.Tensor:
    typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_CudaView(ptr, shape, typestr), device=_require_cuda())


def profile_kernels(fn, max_records: int = 4096):
    """Run fn() with per-launch CUDA-event timing enabled; returns [(kernel name, ms), ...] in launch order."""
    lib().apax_profile_enable(1)
    try:
        fn()
        torch.cuda.synchronize()
        stride = 64
        names = C.create_string_buffer(stride * max_records)
        ms = (C.c_float * max_records)()
        n = lib().apax_profile_collect(C.cast(names, C.c_void_p), stride, C.cast(ms, C.c_void_p), max_records)
        if n < 0:
            check(n, "apax_profile_collect")
        out = []
        for i in range(n):
            out.append((names.raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode(), float(ms[i])))
        return out
    finally:
        lib().apax_profile_enable(0)
