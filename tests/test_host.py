"""Host-side logic: parser, step aliases, window geometry and .npz layout (against values the
reference produced, tests/golden/manifest.json), the C-ABI library's exports, the no-fallback
rule, and the multi-process partition / gather with gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys
from argparse import Namespace

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from dbspm_b200 import _lib, parallel                      # noqa: E402
from dbspm_b200.grid import Grid, get_grid_from_params, read_grid, save_grid, sph2cart  # noqa: E402
from dbspm_b200.input_parser import default_params, parse_params  # noqa: E402
from dbspm_b200.setup_calculation import setup_calc            # noqa: E402

PYRIDINE_IN = """# Example input
LABEL   = pyridine
Z0      = 2.7
ZF      = 4.5
ZREF   = 3.0   # comment
SAMPLEGRID = True
TIPGRID = True 
ALPHA   = 1.08
V       = 42.9071
OPT = False
&VASP
$NPAR = 8
"""


def _params(tmp_path, text=PYRIDINE_IN):
    f = tmp_path / "input.in"
    f.write_text(text)
    return parse_params(str(f))


def test_parser_types_defaults_and_quirks(tmp_path, manifest):
    p = _params(tmp_path)
    want = manifest["geometry"]["pyridine"]["params"]
    for k in ("ALPHA", "V", "Z0", "ZF", "ZREF", "ZPAD", "RPIVOT", "K", "MINMETHOD"):
        assert getattr(p, k) == want[k], k
    assert p.LABEL == "pyridine" and list(p.COMPONENTS) == ["sr", "es", "vdw"]
    assert p.SAMPLEGRID is True
    assert p.TIPGRID is False            # "True " with a trailing blank is not the exact text "True"
    assert p.BBOX is None and str(p.SAMPLE) == "sample" and "NPAR" not in vars(p)
    assert default_params["V"] == 42.91 and default_params["MINMETHOD"] == "Powell"
    q = _params(tmp_path, "BBOX0 = 0 0 1\nBBOXF = 2 3 4\nCOMPONENTS = 1 2\n")
    assert q.BBOX.shape == (2, 3) and q.BBOX[1, 2] == 4.0


def test_step_aliases():
    on = lambda words: {k for k, v in vars(setup_calc(words)).items() if v}
    assert on(["sr"]) == {"sr"} and on(["pauli"]) == {"sr"}
    assert on(["dens"]) == {"sr", "es"} and on(["static"]) == {"sr", "es", "vdw"}
    assert on(["afm"]) == on(["fdbm"]) == {"sr", "es", "vdw", "relax"}
    assert on(["fullauto"]) == {"sample", "tip", "grid", "sr", "es", "vdw", "relax"}
    assert on(["nonsense", "relax"]) == {"relax"}


@pytest.mark.parametrize("name,shape,cell,z0,zf", [
    ("pyridine", (196, 196, 240), (15.0, 15.0, 18.0), 2.7, 4.5),
    ("triangulene", (240, 320, 240), (18.0, 24.0, 18.0), 2.5, 5.0),
])
def test_window_geometry_matches_reference(manifest, name, shape, cell, z0, zf):
    want = manifest["geometry"][name]
    p = Namespace(**{**default_params, "Z0": z0, "ZF": zf, "ZREF": 3.0})
    g = Grid(shape, cell=np.diag(cell), positions=np.zeros((1, 3)), numbers=np.array([1]))
    assert np.allclose(g.dr, want["dr"], rtol=0, atol=0)
    win, nidx = get_grid_from_params(p, g, nidx=True, rpivot=p.RPIVOT, zpad=p.ZPAD)
    assert [int(v) for v in win.shape] == want["window_shape"]
    assert [nidx[2].start, nidx[2].stop] == want["window_z"]
    assert np.array_equal(win.origin, want["window_origin"]) and np.array_equal(win.span, want["window_span"])
    rel, ridx, nbox = get_grid_from_params(p, g, nidx=True, nbox=True)
    assert [int(v) for v in rel.shape] == want["relax_shape"] and [ridx[2].start, ridx[2].stop] == want["relax_z"]
    assert rel.z[0] == want["relax_z_first"] and rel.z[-1] == want["relax_z_last"]
    assert p.BBOX is not None                     # side effect kept from the reference


def test_npz_layout_roundtrip(tmp_path):
    g = Grid((4, 5, 6), cell=np.diag([2.0, 2.5, 3.0]), positions=np.zeros((2, 3)), numbers=np.array([1, 6]))
    data = np.arange(120.0).reshape(4, 5, 6)
    g.set_density("sr", data)
    f = tmp_path / "x_sr_a1.08.npz"
    g.save_density("sr", filename=str(f))
    z = np.load(f)
    assert z.files == ["sr", "shape", "span", "origin", "ortho", "cell", "numbers", "positions", "zref"]
    assert z["sr"].dtype == np.float64 and np.array_equal(z["sr"], data)
    back = read_grid(str(f))
    assert np.array_equal(back.sr, data) and np.array_equal(back.dr, g.dr)
    with pytest.raises(AssertionError):
        save_grid(str(f), g, sr=np.zeros((4, 5, 7)))
    with pytest.raises(ValueError):
        Grid((4, 4, 4), cell=np.array([[1.0, 0, 0.2], [0, 1, 0], [0, 0, 1]]))   # z not orthogonal


def _declared_prototypes():
    """Function names the C compiler sees after preprocessing include/dbspm_b200.h (comments stripped by gcc -E: a
    prototype swallowed by an unterminated comment would be missing here, unlike in a regex over the raw text)."""
    hdr = os.path.join(ROOT, "include", "dbspm_b200.h")
    out = subprocess.run(["gcc", "-E", "-P", "-x", "c", hdr], capture_output=True, text=True, check=True).stdout
    return set(re.findall(r"\b(dbspm_[a-z0-9_]+)\s*\(", out))


def test_header_compiles_as_c_and_cpp_and_declares_every_export():
    hdr = os.path.join(ROOT, "include", "dbspm_b200.h")
    for lang, cc in (("c", "gcc"), ("c++", "g++")):
        r = subprocess.run([cc, "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-x", lang, hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    declared = _declared_prototypes()
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)


_C_CONSUMER = r"""
#include <stdio.h>
#include <string.h>
#include "dbspm_b200.h"
int main(void)
{
    const char *v = dbspm_version();
    if (!v || !strstr(v, "sm_100a")) return 10;
    /* argument validation happens before any CUDA call: usable on a box without a GPU */
    const double dr[3] = {0.1, 0.1, 0.1};
    int rc = dbspm_relax_powell_f64(NULL, 4, 4, 4, 0, 4, 0, 4, 2, 0.2, 3.0, dr, NULL, 1e-4, 1e-4, 2000,
                                    NULL, NULL, NULL, NULL, 0, NULL);
    if (rc != DBSPM_EINVAL) return 11;
    if (!strstr(dbspm_last_error(), "null")) return 12;
    double x = 0.0;
    rc = dbspm_relax_powell_f64(&x, 4, 4, 4, 0, 4, 0, 4, 2, 0.2, 3.0, dr, NULL, 1e-4, 1e-4, 0, &x, NULL, NULL, NULL, 0, NULL);
    if (rc != DBSPM_EINVAL || !strstr(dbspm_last_error(), "maxfev")) return 13;
    if (dbspm_corr3d_workspace_bytes(196, 196, 240, 76, 130, 0) != (size_t)2 * 196 * 196 * 240 * 8) return 14;
    if (dbspm_corr3d_f64(NULL, NULL, 1.0, 1.0, 1.0, 4, 4, 4, 0, 4, NULL, NULL, 0, 0, NULL) != DBSPM_EINVAL) return 15;
    printf("%s\n", v);
    return 0;
}
"""


def test_c_program_links_against_the_library(tmp_path):
    """A 20-line C consumer compiled against include/dbspm_b200.h with -Wall -Werror (an implicit declaration -- a
    prototype missing from the header -- is an error) and linked against libdbspm_b200.so."""
    src = tmp_path / "consumer.c"
    src.write_text(_C_CONSUMER)
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Werror=implicit-function-declaration",
                        "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-ldbspm_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "sm_100a" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_library_loads_and_exports_every_declared_symbol():
    declared = _declared_prototypes()
    L = _lib.lib()                                 # dlopen; no compute call without a GPU
    for name in declared:
        assert getattr(L, name) is not None
    assert "sm_100a" in _lib.version()
    assert L.dbspm_corr3d_workspace_bytes(196, 196, 240, 76, 130, 0) == 2 * 196 * 196 * 240 * 8
    assert L.dbspm_corr3d_workspace_bytes(4, 4, 4, 3, 2, 0) == 0
    # argument validation happens before any CUDA call
    rc = L.dbspm_corr3d_f64(None, None, 1.0, 1.0, 1.0, 4, 4, 4, 0, 4, None, None, 0, 0, None)
    assert rc == _lib.EINVAL and b"null" in L.dbspm_last_error()


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_sass_has_the_async_machinery_the_design_claims():
    """DESIGN.md names TMA tensor loads/stores (axis_tma_kernel), the bulk L2 prefetch of the z kernel and the L2
    prefetch of the y passes; their SASS mnemonics must be in the shipped library, and so must the distributed kernels."""
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = out.stdout
    # round 2: cp.async slots of the pipelined axis pass (LDGSTS + LDGDEPBAR), the 256-bit y-quad loads of the relax stencil,
    # the task counter / progress words of the dynamic relax schedule (ATOMG, acquire load, release store, NANOSLEEP)
    for mnemonic in ("UTMALDG", "UTMASTG", "UBLKPF.L2", "CCTL.E.PF2", "DFMA", "LDGSTS.E.BYPASS.128", "LDGDEPBAR", ".256.CONSTANT",
                     "ATOMG.E.ADD.STRONG.GPU", "LDG.E.STRONG.GPU", "STG.E.STRONG.GPU", "NANOSLEEP"):
        assert mnemonic in sass, mnemonic
    for kernel in ("axis_pass2_kernel", "axis_pass2w_kernel", "axis_pass2m_kernel", "axis_pass2p_kernel", "axis_pass2pw_kernel",
                   "axis_pass2s_kernel", "axis_pass2sw_kernel", "axis_tma_kernel", "zfused2_kernel", "relax_kernel", "relax_dyn_kernel",
                   "direct_kernel"):
        assert kernel in sass, kernel


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dbspm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "liboracle" not in text and "libdbspm_emu" not in text, f


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dbspm_b200 import interactions as ops
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.get_density_overlap(np.zeros((4, 4, 4)), np.zeros((4, 4, 4)), np.ones(3))
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.relax_grid(np.zeros((4, 4, 4)), 0.2, 3.0, np.ones(3), 2)
    with pytest.raises(NotImplementedError):
        ops.relax_grid(np.zeros((4, 4, 4)), 0.2, 3.0, np.ones(3), 2, method="BFGS")

    class HostInterp:
        def ip(self, xyz):
            return 0.0
    with pytest.raises(TypeError):
        ops.relax_tip_z(0, 0, HostInterp(), 0.2, 40.0, np.arange(3), np.ones(3))


def test_split_range():
    assert parallel.split_range(76, 130, 8) == [(76, 82), (82, 89), (89, 96), (96, 103), (103, 109),
                                                (109, 116), (116, 123), (123, 130)]
    assert parallel.split_range(0, 3, 4) == [(0, 0), (0, 1), (1, 2), (2, 3)]
    parts = parallel.split_range(0, 1024, 8)
    assert parts[0] == (0, 128) and parts[-1] == (896, 1024)
    # sr / es interaction plan: 1 rank runs both; N >= 2 ranks split into an sr half and an es half
    assert parallel.interaction_plan(1, 76, 130) == [[(0, 76, 130), (1, 76, 130)]]
    p4 = parallel.interaction_plan(4, 76, 130)
    assert p4 == [[(0, 76, 103)], [(0, 103, 130)], [(1, 76, 103)], [(1, 103, 130)]]
    p3 = parallel.interaction_plan(3, 0, 10)
    assert p3 == [[(0, 0, 10)], [(1, 0, 5)], [(1, 5, 10)]]
    for world in (2, 3, 5, 8):
        cover = {0: [], 1: []}
        for jobs in parallel.interaction_plan(world, 76, 130):
            for w, zl, zh in jobs:
                cover[w].extend(range(zl, zh))
        assert sorted(cover[0]) == sorted(cover[1]) == list(range(76, 130))


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
from dbspm_b200 import parallel
rank, world, device = parallel.init_from_env("gloo")
assert world == 2 and device.type == "cpu"
rng = np.random.default_rng(5)
full = torch.from_numpy(rng.random((5, 4, 7)))            # "window" every rank could compute
def slab(zl, zh):
    return full[:, :, zl - 10:zh - 10].contiguous()
win = parallel.corr3d_sharded(slab, 5, 4, 10, 17, device)
tiles = {{k: torch.from_numpy(rng.random((5, 4, 3))) for k in ("E", "tip_x")}}
def tile(i_lo, i_hi):
    return {{k: v[i_lo:i_hi].contiguous() for k, v in tiles.items()}}
res = parallel.relax_sharded(tile, 5, device)
# no symmetric memory on the CPU: the peer assembly declines and callers fall back to the gathers below
assert parallel.PeerAssembler.get("t", (6, 4, 3), device) is None
# equal tiles are received straight into place (no padded copies, no concatenation)
tiles6 = {{k: torch.from_numpy(rng.random((6, 4, 3))) for k in ("E", "tip_x")}}
res6 = parallel.relax_sharded(lambda a, b: {{k: v[a:b] for k, v in tiles6.items()}}, 6, device)
assert (res6 is None) == (rank != 0)
if rank == 0:
    assert all(torch.equal(res6[k], tiles6[k]) for k in tiles6), "x-tile gather in place"
# sr on rank 0, es on rank 1 (interaction-parallel plan), one gather assembles both windows
plan = parallel.interaction_plan(2, 10, 17)
assert plan == [[(0, 10, 17)], [(1, 10, 17)]]
full2 = [full, full * 2.0 + 1.0]
which, zl, zh = plan[rank][0]
wins = parallel.gather_windows(plan, [full2[which][:, :, zl - 10:zh - 10].contiguous()], 5, 4, 10, 17, device)
if rank == 0:
    assert torch.equal(wins[0], full2[0]) and torch.equal(wins[1], full2[1]), "interaction gather"
else:
    assert wins is None
# parameter sweep: sets dealt by alpha, results back on rank 0 in the caller's order
sets = [(1.08, 40.0, 0.2), (1.0, 40.0, 0.2), (1.08, 50.0, 0.2), (1.0, 50.0, 0.24), (1.12, 35.0, 0.2)]
asg = parallel.sweep_assignment(sets, 2)
assert asg == [[1, 3], [0, 2, 4]], asg
mine = [{{"E": torch.full((2, 3), float(i)), "tip_x": torch.full((2, 3), 10.0 + i)}} for i in asg[rank]]
got = parallel.gather_sweep(mine, asg, keys=("E", "tip_x"))
if rank == 0:
    assert [float(g["E"][0, 0]) for g in got] == [0.0, 1.0, 2.0, 3.0, 4.0]
    assert [float(g["tip_x"][1, 2]) for g in got] == [10.0, 11.0, 12.0, 13.0, 14.0]
else:
    assert got is None
# distributed transform: one interaction over both ranks (x-slabs, all-to-all in slot order, gather, final y pass)
class NumpyStages:
    # the three compute stages of dbspm_b200.interactions.dist_* restated with numpy on unpacked complex arrays
    @staticmethod
    def dist_forward(a, b, G, alpha0, alpha1):
        out = []
        for arr, al in ((a, alpha0), (b, alpha1)):
            f = np.fft.fft(arr.numpy() ** al, axis=1)[:, parallel.slot_order(arr.shape[1]), :]
            nxl, ny, nz = f.shape
            blk = np.ascontiguousarray(f.reshape(nxl, G, ny // G, nz).transpose(1, 0, 2, 3))
            out.append(torch.from_numpy(blk.view(np.float64)))
        return tuple(out)
    @staticmethod
    def dist_middle(ra, rb, dr, shape, G, h, z_lo, z_hi):
        nx, ny, nz = shape
        A = ra.numpy().view(np.complex128).reshape(nx, ny // G, nz)
        B = rb.numpy().view(np.complex128).reshape(nx, ny // G, nz)
        S = np.fft.fft(np.fft.fft(A, axis=0), axis=2) * np.conj(np.fft.fft(np.fft.fft(B, axis=0), axis=2)) * np.prod(dr)
        w = np.fft.ifft(np.fft.ifft(S, axis=2)[:, :, z_lo:z_hi], axis=0)
        return torch.from_numpy(np.ascontiguousarray(w).view(np.float64))
    @staticmethod
    def dist_finish(W_all, shape, G, z_lo, z_hi):
        nx, ny, nz = shape
        w = W_all.numpy().view(np.complex128)                              # (G, nx, ny/G, nzw)
        slots = w.transpose(1, 0, 2, 3).reshape(nx, ny, z_hi - z_lo)
        nat = np.empty_like(slots)
        nat[:, parallel.slot_order(ny), :] = slots
        return torch.from_numpy(np.fft.ifft(nat, axis=1).real.copy())
shape = (6, 8, 10)
A = rng.random(shape); B = rng.random(shape) + 0.1
drr = np.array([0.1, 0.2, 0.3])
x0, x1 = parallel.x_slab(shape[0], 2, rank)
wins = parallel.sr_es_distributed(NumpyStages, torch.from_numpy(A[x0:x1].copy()), torch.from_numpy(B[x0:x1].copy()), 1.08,
                                  shape, drr, 2, 9, [[0, 1]])
ref_all = (np.fft.ifftn(np.fft.fftn(A ** 1.08) * np.conj(np.fft.fftn(B ** 1.08))).real * np.prod(drr))[:, :, 2:9]
if rank == 0:
    ref = ref_all
    assert np.abs(wins[0].numpy() - ref).max() <= 1e-12 * np.abs(ref).max(), "distributed transform"
else:
    assert wins is None
# the multi-GPU step of round 2: every rank keeps its finished x-slab (no gather), builds its static tile, fetches a halo
# of the lever length from its neighbours (periodic; with two ranks both neighbours are the same peer), and only the
# .npz needs the window on rank 0 (asynchronous gather of the slabs)
slab = parallel.sr_es_distributed(NumpyStages, torch.from_numpy(A[x0:x1].copy()), torch.from_numpy(B[x0:x1].copy()), 1.08,
                                  shape, drr, 2, 9, [[0, 1]], gather=False)
assert tuple(slab.shape) == (3, 8, 7) and np.abs(slab.numpy() - ref_all[x0:x1]).max() <= 1e-12 * np.abs(ref_all).max()
for halo in (1, 2, 3, 5):                                   # 5 > slab: the all-gather fallback
    loc = parallel.exchange_halo(slab, halo)
    want = ref_all[(np.arange(x0 - halo, x1 + halo)) % shape[0]]
    assert tuple(loc.shape) == (3 + 2 * halo, 8, 7) and np.abs(loc.numpy() - want).max() <= 1e-12 * np.abs(ref_all).max(), halo
assert parallel.exchange_halo(slab, 0) is slab
winx, work = parallel.gather_x_slabs(slab, shape[0], async_op=True)
work.wait()
if rank == 0:
    assert np.abs(winx.numpy() - ref_all).max() <= 1e-12 * np.abs(ref_all).max()
else:
    assert winx is None
assert parallel.dist_groups(8) == [[0, 1, 2, 3], [4, 5, 6, 7]] and parallel.dist_groups(2) is None
assert parallel.dist_groups(6) == [[0, 1, 2], [3, 4, 5]] and parallel.dist_groups(5) is None
t = torch.arange(3.0) if rank == 0 else torch.zeros(3)
parallel.broadcast_(t)
assert torch.equal(t, torch.arange(3.0))
if rank == 0:
    assert torch.equal(win, full), "z-slab gather/interleave"
    assert all(torch.equal(res[k], tiles[k]) for k in tiles), "x-tile gather"
    print("RANK0_OK")
else:
    assert win is None and res is None
torch.distributed.barrier()
torch.distributed.destroy_process_group()
'''


def test_partition_and_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "RANK0_OK" in outs[0]


def test_sph2cart_matches_definition():
    th, az = np.array([3.0, 2.5]), np.array([0.3, -1.0])
    x, y, z = sph2cart(th, az, 3.02)
    assert np.allclose(x ** 2 + y ** 2 + z ** 2, 3.02 ** 2)
    assert np.allclose(z, 3.02 * np.cos(th))


def test_circular_support_host_logic():
    from dbspm_b200.interactions import circular_support
    f = np.zeros(40, bool); f[[0, 1, 2, 37, 38, 39]] = True
    assert circular_support(f) == (37, 6)
    f = np.zeros(40, bool); f[[5, 9]] = True
    assert circular_support(f) == (5, 5)
    assert circular_support(np.zeros(8, bool)) == (0, 1)
    assert circular_support(np.ones(8, bool)) == (0, 8)
    f = np.zeros(10, bool); f[[9]] = True
    assert circular_support(f) == (9, 1)


def test_sweep_assignment_balances_and_groups_alphas():
    from dbspm_b200.parallel import sweep_assignment
    sets = [(a, v, 0.2) for a in (1.0, 1.04, 1.08, 1.12) for v in (35.0, 42.9, 50.0, 60.0)]
    for world in (1, 2, 3, 4, 8, 16, 20):
        asg = sweep_assignment(sets, world)
        assert sorted(i for a in asg for i in a) == list(range(16))
        assert max(len(a) for a in asg) - min(len(a) for a in asg) <= 1
    asg = sweep_assignment(sets, 4)
    assert all(len({sets[i][0] for i in a}) == 1 for a in asg)       # one sr grid per rank


def test_npz_reader_matches_numpy_load(tmp_path):
    """dbspm_b200.npzio.read_npz (direct payload reads, pinned when a GPU is present) returns what np.load
    returns, for stored and compressed files, big and small members, and object members."""
    import torch
    from dbspm_b200.npzio import read_npz
    rng = np.random.default_rng(3)
    a = rng.random((96, 96, 64)); b = (rng.random((300, 300, 8)) * 10).astype(np.int32)
    f = tmp_path / "sample.npz"
    np.savez(f, rhoS=a, potS=-a, big=b, cell=np.eye(3), numbers=np.array([1, 6]), zref=np.array(None, dtype=object))
    r = read_npz(f)
    with np.load(f, allow_pickle=True) as ref:
        assert set(r) == set(ref.files)
        for k in ref.files:
            got = r[k].numpy() if isinstance(r[k], torch.Tensor) else r[k]
            assert got.dtype == ref[k].dtype and got.shape == ref[k].shape
            assert (got == ref[k]).all() if got.dtype != object else got.item() is None
    assert isinstance(r["rhoS"], torch.Tensor) and isinstance(r["cell"], np.ndarray)
    assert list(read_npz(f, keys=["potS"])) == ["potS"]
    with pytest.raises(KeyError):
        read_npz(f, keys=["nope"])
    fc = tmp_path / "c.npz"
    np.savez_compressed(fc, rhoS=a)
    assert np.array_equal(read_npz(fc)["rhoS"], a)
