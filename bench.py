"""Benchmark of the importance-nested-sampling cycle (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one pass of the hot path over one batch of 2^20 raw proposals per
GPU drawn from the config-2 bound (30-D Gaussian, n_live=2000, 4 networks
(100, 50, 20); the bound was built by the reference and is shipped as
tests/golden/cfg2_bound_d30.npz): union proposal -> unit-cube filter ->
overlap acceptance -> neural filter -> likelihood -> log-sum-exp / ESS /
counters, and for N > 1 the one NCCL collective that merges the per-rank
sums.  `value` = raw proposals per second over all ranks.

Prints ONE JSON line on rank 0 (see the contract in the task description).
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CPU_BUDGET_S = 60.0      # CPU loop time the reference arm may spend

# BASELINE.json configurations with a fixed, reference-built bound.  Config 2
# is the headline (the driver's default run); config 5 is the strong-scaling
# workload (fixed global batch split over the GPUs).
CONFIGS = {
    2: dict(d=30, golden='cfg2_bound_d30', like='Gaussian', batch=1 << 20,
            scaling='weak', steps=1000,
            workload='cfg2: 30-D isotropic Gaussian (sigma=0.1), '
                     'n_live=2000, bound built by the reference (K=1 '
                     'ellipsoid, 4 nets 100-50-20), batch=2^20 raw '
                     'proposals per GPU per step'),
    5: dict(d=100, golden='cfg5_bound_d100', like='EquicorrelatedGaussian',
            batch=1 << 24, scaling='strong', steps=20,
            workload='cfg5: 100-D correlated Gaussian (sigma=0.05, rho=0.5), '
                     'n_live=10000, bound built by the reference from 40000 '
                     'target draws (4 nets 100-50-20), batch=2^24 raw '
                     'proposals per step IN TOTAL, split over the GPUs'),
}
CONFIG = 2
D = 30
ALGO_BYTES_PER_PROPOSAL = 8 * D + 8 + 1      # row + log_l + disposition
MLP_FLOPS_PER_POINT = 4 * 2 * (30 * 100 + 100 * 50 + 50 * 20 + 20 * 1)
WORKLOAD = CONFIGS[2]['workload']


def select_config(c):
    global CONFIG, D, ALGO_BYTES_PER_PROPOSAL, MLP_FLOPS_PER_POINT, WORKLOAD
    global _SPEC
    cfg = CONFIGS[c]
    CONFIG, D, WORKLOAD = c, cfg['d'], cfg['workload']
    ALGO_BYTES_PER_PROPOSAL = 8 * D + 8 + 1
    MLP_FLOPS_PER_POINT = 4 * 2 * (D * 100 + 100 * 50 + 50 * 20 + 20 * 1)
    _SPEC = None


def make_like():
    from nautilus_b200 import likelihoods
    return getattr(likelihoods, CONFIGS[CONFIG]['like'])(D)


def workload_config(batch):
    """The `config` object: identical for both arms (same workload)."""
    return {'workload': WORKLOAD, 'batch_per_gpu': batch,
            'l2': 'GPU arm: each step writes {:.0f} MB of proposals (> 126 MB '
                  'L2), no flush needed; bound parameters (~520 KB) are meant '
                  'to stay cache-resident'.format(
                      batch * ALGO_BYTES_PER_PROPOSAL / 1e6)}


_SPEC = None


def load_spec():
    global _SPEC
    if _SPEC is None:
        from nautilus_b200._pack import flat_to_spec
        with np.load(os.path.join(ROOT, 'tests', 'golden',
                                  CONFIGS[CONFIG]['golden'] + '.npz')) as f:
            _SPEC = flat_to_spec({k: f[k] for k in f.files})
    return _SPEC


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p['hbm_gbs']), bf16=float(p['bf16_tflops']),
                    bf16_sustained=float(p['bf16_tflops_sustained']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0,
                source='fallback (B200_PROFILING.md)')


# --------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference loop, on host cores
# --------------------------------------------------------------------------

REF_N_BATCH = 1000      # n_batch of the reference Sampler in the CPU arm


def cpu_kind():
    """'reference' when the unmodified reference travelled with the snapshot
    (oracle/_ref, placed by oracle/make_ref.sh), else the oracle port."""
    from oracle import ref_arm
    return 'reference' if ref_arm.available() else 'port'


def _log_l_min():
    # (the config-2 fixture stores the threshold without the normalisation)
    norm = make_like().norm if CONFIG == 2 else 0.0
    with np.load(os.path.join(ROOT, 'tests', 'golden',
                              CONFIGS[CONFIG]['golden'] + '.npz')) as f:
        return float(f['log_l_min']) + norm


def _cpu_worker(args):
    """One host core's share of the CPU arm, single-threaded BLAS like the
    reference (sampler.py:789).

    kind 'reference': the UNMODIFIED reference (oracle/_ref/nautilus): its own
    ``Sampler.add_samples`` (sampler.py:1093-1144) on the config-2 bound until
    the outer union has consumed ~n_raw raw proposals (oracle/ref_arm.py).
    kind 'port': the oracle's restatement of the same loop
    (NautilusBound.sample + likelihood + update_shell_info; 1000 raw draws per
    iteration, bounds/union.py:305-323, bounds/nautilus.py:213-222,
    sampler.py:925-943)."""
    seed, n_raw, kind = args
    from threadpoolctl import threadpool_limits
    spec = load_spec()
    like = make_like()
    if kind == 'reference':
        from oracle import ref_arm
        ref_arm._import_reference()          # not part of the timed loop
        with threadpool_limits(limits=1):
            n, dt, _ = ref_arm.run_reference_cycles(
                spec, like, _log_l_min(), n_raw, seed=seed,
                n_batch=REF_N_BATCH)
        return n, dt
    from oracle import nautilus_oracle as orc
    rng = np.random.default_rng(seed)
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        state = None
        log_l = []
        n_points = 0
        while state is None or state['u_n_sample'] < n_raw:
            pts, state = orc.nautilus_sample(spec, rng, 100, state)
            log_l.append(like(pts))
            n_points += len(pts)
        log_l = np.concatenate(log_l)
        log_v = orc.bound_log_v(spec, state['u_n_sample'],
                                state['u_n_reject'], state['n_sample'],
                                state['n_reject'])
        orc.shell_info(log_l, log_v, n_points)
        dt = time.perf_counter() - t0
    return state['u_n_sample'], dt


class CpuPool:
    """Persistent worker processes, one per host core, each running its own
    replica of the CPU loop (forked once, reused for every step)."""

    def __init__(self, cores, kind=None):
        import multiprocessing as mp
        self.cores = cores
        self.kind = kind or cpu_kind()
        self.pool = mp.get_context('fork').Pool(cores) if cores > 1 else None

    def run(self, n_raw_per_core, seed=0):
        """One pass of the CPU loop on every core; returns (raw proposals,
        seconds) with seconds = the slowest worker's loop time."""
        jobs = [(seed + i, n_raw_per_core, self.kind)
                for i in range(self.cores)]
        if self.pool is None:
            res = [_cpu_worker(jobs[0])]
        else:
            res = self.pool.map(_cpu_worker, jobs, chunksize=1)
        return sum(r[0] for r in res), max(r[1] for r in res)

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def cpu_pass(cores, n_raw_per_core, seed=0, kind=None):
    """Run the CPU loop on `cores` processes; returns (raw proposals, s)."""
    pool = CpuPool(cores, kind)
    try:
        return pool.run(n_raw_per_core, seed)
    finally:
        pool.close()


def cpu_sample_text(kind, steps, cores, per_core):
    if kind == 'reference':
        return ('{} step(s) x {} processes x ~{} raw proposals each through '
                'the UNMODIFIED reference (oracle/_ref/nautilus v1.0.6): '
                'Sampler.add_samples(n_batch={}) -> NautilusBound.sample -> '
                'Union.sample -> NeuralBound.contains (scikit-learn predict) '
                '-> likelihood -> update_shell_info on the config-2 bound; '
                'BLAS pinned to 1 thread per process as the reference does'
                .format(steps, cores, per_core, REF_N_BATCH))
    return ('{} step(s) x {} processes x {} raw proposals through the oracle '
            'port of Union.sample/NautilusBound.sample/likelihood/'
            'update_shell_info (oracle/_ref missing on this box)'
            .format(steps, cores, per_core))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path
    on all host cores, bounded samples."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_cores()
    kind = cpu_kind()
    # bounded sample: about CPU_BUDGET_S seconds of loop time for the whole
    # run whatever --steps is
    rate = 6.0e4 if kind == 'reference' else 2.0e5     # raw/s on one core
    per_core = int(CPU_BUDGET_S * rate / max(args.steps, 1))
    per_core = max(2000, min(400000, per_core // 1000 * 1000))
    pool = CpuPool(cores, kind)
    for _ in range(min(args.warmup, 3)):
        pool.run(2000)
    total, wall = 0, 0.0
    for s in range(args.steps):
        n, dt = pool.run(per_core, seed=1000 * (s + 1))
        total += n
        wall += dt
    pool.close()
    value = total / wall
    sample = cpu_sample_text(kind, args.steps, cores, per_core)
    line = {
        'impl': 'reference', 'metric': 'raw_proposals_per_sec',
        'value': value, 'unit': 'proposals/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / max(args.steps, 1),
        'higher_is_better': True, 'scaling': CONFIGS[CONFIG]['scaling'],
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.batch),
        'notes': {'emulator_arith': 'f64 (scikit-learn)',
                  'parallelism': '{} independent host processes'.format(
                      cores)},
        'cpu_baseline': {'value': value, 'unit': 'proposals/s',
                         'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'proposals/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------

class ClockSampler:
    """One `nvidia-smi -lms 50` process for the duration of the timed region
    (the recipe of B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,'
         'clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.rows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        for ln in out.strip().splitlines():
            self.rows.append([c.strip() for c in ln.split(',')])

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [],
                    'samples': 0}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx),
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------

def run_gpu(args):
    import torch
    import torch.distributed as dist
    from nautilus_b200 import likelihoods, ops
    from nautilus_b200.pool import (exchange_packed_async, merge_packed,
                                    packed_stats)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run for --gpus > 1')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # keep stdout for the one JSON line: NCCL's version banner (printed when
        # the box sets NCCL_DEBUG) goes to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)

    spec = load_spec()
    like = make_like()
    stack = ops.DeviceStack([spec], device=dev)
    like_params = like.device_params(dev)
    strong = CONFIGS[CONFIG]['scaling'] == 'strong'
    # weak scaling: args.batch per GPU; strong: args.batch in total
    n = args.batch // world if strong else args.batch
    mode = {'tf32': ops.MLP_TF32, 'f16': ops.MLP_F16,
            'f64': ops.MLP_F64}[args.mlp]
    seed = 0
    log_l_min = _log_l_min()

    out = stack.cycle(0, n, seed=seed, like_id=like.like_id,
                      like_params=like_params, log_l_min=log_l_min, mode=mode)
    # The kernels write the sums and counters straight into the send buffer
    # of the one collective.  Send buffers are double-buffered and the
    # all-gather of step s is asynchronous (NCCL's stream, behind the kernels
    # of step s): the compute stream goes straight on to step s + 1 and only
    # waits for the gather of step s when step s + 2 is about to overwrite its
    # send buffer -- by then it finished long ago, so the exchange is off the
    # critical path.
    outs, words, gathered, works = [], [], [], [None, None]
    for b in range(2):
        w, lse_b, cnt_b = packed_stats(dev, ops.N_CNT, ops.N_LSE)
        o = dict(out)
        o['lse'], o['counters'] = lse_b, cnt_b
        outs.append(o)
        words.append(w)
        gathered.append(torch.zeros((world, ops.N_LSE + ops.N_CNT),
                                    dtype=torch.int64, device=dev))
    state = {'step': 0, 'merged': None}

    def step():
        # every rank draws its own slice of the global proposal index space
        s = state['step']
        state['step'] += 1
        b = s & 1
        if works[b] is not None:
            works[b].wait()          # send buffer b is free again
            works[b] = None
        offset = (s * world + rank) * n
        stack.cycle(0, n, seed=seed, offset=offset, like_id=like.like_id,
                    like_params=like_params, log_l_min=log_l_min, mode=mode,
                    out=outs[b])
        if world > 1:
            # the one exchange step: per-rank counters + LSE partials; like
            # the single-GPU results they stay on the device until read
            works[b] = exchange_packed_async(words[b], gathered=gathered[b],
                                             async_op=True)

    def drain():
        for b in range(2):
            if works[b] is not None:
                works[b].wait()
                works[b] = None

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.3)

    # ---- timed region: K steps, device time, max over ranks ---------------
    launches0 = ops.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    drain()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if clocks:
        clocks.stop()

    # result of the last step (also a sanity check on the collective)
    last = (state['step'] - 1) & 1
    if world > 1:
        cnt, (m, s1, s2) = merge_packed(gathered[last], ops.N_LSE)
        cnt = cnt.astype(float)
    else:
        cnt = outs[last]['counters'].cpu().numpy().astype(float)
        m, s1, s2 = outs[last]['lse'].cpu().numpy()[:3]
    n_shell = cnt[ops.CNT_IN_SHELL]
    result = {
        'raw': int(cnt[ops.CNT_RAW]), 'in_shell': int(n_shell),
        'acceptance': float(n_shell / cnt[ops.CNT_RAW]),
        'shell_log_l': float(m + np.log(s1) - np.log(n_shell)),
        'shell_n_eff': float(s1 * s1 / s2),
    }

    # ---- roofline leg: same K steps with per-stage events ------------------
    barrier()
    ops.profile_enable(True)
    for _ in range(args.steps):
        step()
    prof = ops.profile_collect()
    ops.profile_enable(False)
    barrier()

    # ---- e2e leg: the C-ABI host-buffer session (include/nautilus_b200.h,
    # nb200_session_*): NumPy in, NumPy out.  Every step uploads the
    # serialised bound + likelihood parameters from pinned host memory and
    # brings back what Sampler.add_samples keeps of the batch
    # (sampler.py:1135-1141) -- for every in-shell proposal its GLOBAL index
    # (u64, from which nb200_session_materialize regenerates the row bit for
    # bit) and its log_l, plus the counters and the log-sum-exp triple; the
    # rows themselves stay on the device.  Two batches are in flight so the
    # device->host copy of one runs under the kernels of the next.  The
    # round-1 form (rows cross PCIe, 8d+8 bytes per in-shell proposal) is
    # timed as well, on fewer steps, and reported as e2e.rows_mode.
    par_h = like_params.cpu().numpy()

    def e2e_leg(returns, k_steps):
        sess = ops.HostSession([spec], n_max=n, cap=max(n // 4, 1024),
                               n_slots=2, returns=returns,
                               like_params_max=like_params.numel())
        io = {'h2d': 0, 'd2h': 0, 'points': 0, 'sum_ll': 0.0, 'last': None}

        def submit(slot):
            s = state['step']
            state['step'] += 1
            sess.submit(slot, 0, n, seed=seed, offset=(s * world + rank) * n,
                        like_id=like.like_id, like_params=par_h,
                        log_l_min=log_l_min, mode=mode, upload_stack=True)

        def wait(slot):
            res = sess.wait(slot)
            k = len(res['log_l'])
            io['points'] += k
            io['sum_ll'] += float(res['log_l'][-1]) if k else 0.0
            io['h2d'] = sess.stack_bytes + par_h.nbytes
            small = (ops.N_LSE + ops.N_CNT + 1) * 8
            if returns == 'index':
                # submit enqueues the copy before the count is known: a
                # quarter more than the previous batch kept (+1024) is sent
                io['d2h'] = (k + k // 4 + 1024) * 16 + small
                io['last'] = np.array(res['index'][:64])
            else:
                io['d2h'] = k * (D + 1) * 8 + small

        def run(steps):
            submit(0)
            for i in range(steps):
                if i + 1 < steps:
                    submit((i + 1) & 1)
                wait(i & 1)

        run(3)
        barrier()
        t0 = time.perf_counter()
        run(k_steps)
        ms_leg = 1e3 * (time.perf_counter() - t0)
        barrier()
        if returns == 'index' and io['last'] is not None:
            # rows on demand (outside the timed region): regenerate a few
            rows = sess.materialize(0, io['last'], seed=seed, mode=mode)
            assert rows.shape == (len(io['last']), D) and np.all(
                (rows >= 0) & (rows < 1))
        sess.close()
        if world > 1:
            t = torch.tensor([ms_leg], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_leg = float(t.item())
        return ms_leg, io

    e2e_ms, bytes_io = e2e_leg('index', args.steps)
    rows_steps = max(3, min(args.steps, 100))
    rows_ms, rows_io = e2e_leg('rows', rows_steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    value = world * n * args.steps / (ms * 1e-3)
    e2e_value = world * n * args.steps / (e2e_ms * 1e-3)
    stage_ms = {k: v[0] / args.steps for k, v in prof.items() if v[1] > 0}
    total_stage = sum(stage_ms.values()) or 1.0
    top = max(stage_ms, key=stage_ms.get)
    evaluated = result['raw'] / world - 0   # ~all proposals reach the MLP
    traffic, traffic_all = None, None
    tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_all = json.load(f)
        traffic = traffic_all.get(top)
    if top == 'mlp_predict':
        flops = MLP_FLOPS_PER_POINT * evaluated
        achieved = flops / (stage_ms[top] * 1e-3) / 1e12
        roofline = {
            'kernel': top, 'bound': 'tensor', 'achieved': achieved,
            'peak': peaks['bf16'], 'unit': 'TFLOP/s',
            'frac': achieved / peaks['bf16'], 'traffic': traffic,
            'peak_source': peaks['source'] + ', dense bf16 burst',
            'flops_per_launch': flops, 'ms_per_launch': stage_ms[top],
            'arith': args.mlp}
    else:
        nbytes = ALGO_BYTES_PER_PROPOSAL * n
        achieved = nbytes / (stage_ms[top] * 1e-3) / 1e9
        roofline = {
            'kernel': top, 'bound': 'hbm', 'achieved': achieved,
            'peak': peaks['hbm'], 'unit': 'GB/s',
            'frac': achieved / peaks['hbm'], 'traffic': traffic,
            'peak_source': peaks['source'],
            'bytes_per_launch': nbytes, 'ms_per_launch': stage_ms[top]}
    roofline['share_of_step'] = stage_ms[top] / total_stage
    roofline['stages_ms'] = stage_ms
    # the other ceilings of the two hot kernels: the
    # front kernel's two triangular matrix-vector products on the FP64 tensor
    # cores (peak measured by tools/peak_dmma.cu) and the emulator's dense
    # contraction on tcgen05 (kind::tf32, whose peak is half the bf16 one)
    fp64_peak = 37.0
    ppath = os.path.join(ROOT, 'profiles', 'r1_peak_dmma.json')
    if os.path.exists(ppath):
        with open(ppath) as f:
            fp64_peak = float(json.load(f)['dmma_tflops_w8_16warps'])
    if 'fused_cycle' in stage_ms:
        # x = B z + c on the FP64 tensor cores.  The second product,
        # t = B_inv (x - c), is skipped by the whitening shortcut whenever the
        # neural bound's ellipsoid is the mixture's (this workload; it is
        # redone only inside the guard band around r^2 = 1), so one
        # triangular product per proposal is what the kernel executes
        fl = 2.0 * (D * (D + 1) / 2) * n
        ach = fl / (stage_ms['fused_cycle'] * 1e-3) / 1e12
        roofline['front_fp64'] = {
            'achieved': ach, 'peak': fp64_peak, 'unit': 'TFLOP/s',
            'frac': ach / fp64_peak, 'flops_per_launch': fl,
            'peak_source': 'measured DMMA m8n8k4 (profiles/r1_peak_dmma.json)'}
        # the front kernel writes every algorithmic byte of the step
        # (row + log_l + disposition): its own HBM figure, whichever kernel
        # is the longest of the step
        nbytes = ALGO_BYTES_PER_PROPOSAL * n
        ach = nbytes / (stage_ms['fused_cycle'] * 1e-3) / 1e9
        roofline['front_hbm'] = {
            'achieved': ach, 'peak': peaks['hbm'], 'unit': 'GB/s',
            'frac': ach / peaks['hbm'], 'bytes_per_launch': nbytes,
            'ms_per_launch': stage_ms['fused_cycle'],
            'traffic': (traffic_all or {}).get('fused_cycle')}
    if 'mlp_predict' in stage_ms:
        fl = MLP_FLOPS_PER_POINT * evaluated
        ach = fl / (stage_ms['mlp_predict'] * 1e-3) / 1e12
        # kind::f16 runs at the bf16 rate; kind::tf32 at half of it
        tpeak = peaks['bf16'] * (0.5 if args.mlp == 'tf32' else 1.0)
        roofline['emulator_tensor'] = {
            'achieved': ach, 'peak': tpeak, 'unit': 'TFLOP/s',
            'frac': ach / tpeak, 'flops_per_launch': fl,
            'arith': args.mlp,
            'peak_source': peaks['source'] + ', dense bf16 burst' + (
                ' / 2 (tf32)' if args.mlp == 'tf32' else
                ' (fp16 operands run at the bf16 rate)')}
        # the instruction's own ceiling in the shape the kernel issues
        # (cta_group::1, M = 128), measured by tools/peak_umma.cu on this
        # pool's B200 and committed under profiles/
        try:
            with open(os.path.join(ROOT, 'profiles', 'r2_peak_umma.json')) as f:
                umma = json.load(f)
            upeak = umma['tf32_ts_n128' if args.mlp == 'tf32'
                         else 'f16_ts_n128']
            roofline['emulator_tensor']['umma_peak'] = upeak
            roofline['emulator_tensor']['frac_of_umma_peak'] = ach / upeak
            roofline['emulator_tensor']['umma_peak_source'] = (
                'profiles/r2_peak_umma.json (tools/peak_umma.cu: tcgen05.mma '
                'cta_group::1 M=128 N=128, A from TMEM, one CTA per SM)')
        except (OSError, KeyError, ValueError):
            pass
    # whole-step figure against the 8d+9 B/proposal HBM roofline
    roofline['cycle_hbm_frac'] = (value / world * ALGO_BYTES_PER_PROPOSAL /
                                  1e9) / peaks['hbm']

    cpu = None
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        kind = cpu_kind()
        per_core = 100000 if kind == 'reference' else 200000
        n_cpu, wall = cpu_pass(cores, per_core, kind=kind)
        n1, wall1 = cpu_pass(1, per_core, kind=kind)
        cpu = {'value': n_cpu / wall, 'unit': 'proposals/s', 'cores': cores,
               'kind': kind,
               'sample': cpu_sample_text(kind, 1, cores, per_core),
               'single_core_value': n1 / wall1}
        if kind == 'reference':
            # cross-check: the oracle port of the same loop on one core
            np_, wp = cpu_pass(1, 200000, kind='port')
            cpu['port_single_core_value'] = np_ / wp

    # ---- second workload: the same cycle in the SAMPLING phase, where a
    # proposal of shell i is also tested against every later bound
    # (sampler.py:796-801; 70 % of the reference's time there).  47 later
    # bounds = nested copies of the config-2 bound (radius factor 0.9885 per
    # bound: the volume halves every two bounds), each with its own threshold.
    later_line = None
    if CONFIG == 2 and not args.no_later:
        import copy
        os.environ.pop('NB200_EXCLUDE', None)

        def nested(f, dthr):
            sp = copy.deepcopy(spec)
            for mx in sp['mixtures']:
                mx['ell']['B'] = mx['ell']['B'] * f
                mx['ell']['B_inv'] = mx['ell']['B_inv'] / f
            for nbs in sp['neural']:
                nbs['ell']['B'] = nbs['ell']['B'] * f
                nbs['ell']['B_inv'] = nbs['ell']['B_inv'] / f
                nbs['score_predict_min'] += dthr
            return sp

        L = 47
        rs = np.random.default_rng(47)
        later = [nested(0.9885**(i + 1), 0.01 * rs.normal())
                 for i in range(L)]
        stack_l = ops.DeviceStack([spec] + later, device=dev)
        out_l = stack_l.cycle(0, n, later=(1, L), seed=seed,
                              like_id=like.like_id, like_params=like_params,
                              log_l_min=log_l_min, mode=mode)

        def later_leg(k_steps):
            for _ in range(2):
                stack_l.cycle(0, n, later=(1, L), seed=seed,
                              like_id=like.like_id, like_params=like_params,
                              log_l_min=log_l_min, mode=mode, out=out_l)
            barrier()
            l0 = ops.launch_count()
            a0 = torch.cuda.Event(enable_timing=True)
            a1 = torch.cuda.Event(enable_timing=True)
            a0.record()
            for s_ in range(k_steps):
                stack_l.cycle(0, n, later=(1, L), seed=seed,
                              offset=(s_ * world + rank) * n,
                              like_id=like.like_id, like_params=like_params,
                              log_l_min=log_l_min, mode=mode, out=out_l)
            a1.record()
            barrier()
            return a0.elapsed_time(a1) / k_steps, \
                (ops.launch_count() - l0) / k_steps

        ms_g, launches_g = later_leg(50)
        cnt_l = out_l['counters'].cpu().numpy()
        os.environ['NB200_EXCLUDE'] = 'loop'
        ms_loop, launches_loop = later_leg(5)
        os.environ.pop('NB200_EXCLUDE', None)
        if world > 1:
            t = torch.tensor([ms_g, ms_loop], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_g, ms_loop = float(t[0].item()), float(t[1].item())
        later_line = {
            'workload': 'cfg2 + 47 later bounds (sampling phase): every '
                        'proposal that passes the bound is tested against 47 '
                        'nested later bounds (47 x 4 networks)',
            'value': world * n / (ms_g * 1e-3), 'unit': 'proposals/s',
            'ms_per_step': ms_g, 'launches_per_step': launches_g,
            'excluded_fraction_of_accepted': float(
                cnt_l[ops.CNT_EXCLUDED] / max(1, cnt_l[ops.CNT_EXCLUDED] +
                                              cnt_l[ops.CNT_IN_SHELL])),
            'per_bound_loop': {'value': world * n / (ms_loop * 1e-3),
                               'ms_per_step': ms_loop,
                               'launches_per_step': launches_loop}}
        del stack_l, out_l

    # |delta log Z| of BASELINE's metric: config 2 end to end through the
    # drop-in Sampler (outside every timed region), against the analytic
    # evidence of the 30-D Gaussian; N_eff >= 4e4 puts the statistical error
    # (1/sqrt(N_eff) = 0.005) well inside north_star's 0.01
    logz = None
    if world == 1 and not args.no_logz and CONFIG == 2:
        from nautilus_b200 import Sampler
        t0 = time.perf_counter()
        smp = Sampler(lambda x: x, like, n_dim=D, n_live=2000, seed=0,
                      emulator_arith=args.mlp)
        # N_eff = 1e5: statistical error 1 / sqrt(N_eff) = 0.0032, so that the
        # 0.01 bar is a 3-sigma statement about the method, not a coin flip
        ok = smp.run(n_eff=100000, discard_exploration=True, timeout=600)
        raw_total = sum(b.outer_bound.n_sample for b in smp.bounds[1:])
        logz = {'log_z': float(smp.log_z), 'log_z_true': like.log_z_true,
                'delta_log_z': abs(float(smp.log_z) - like.log_z_true),
                'target': 0.01, 'n_eff': float(smp.n_eff),
                'stat_error': float(1 / np.sqrt(smp.n_eff)),
                'converged': bool(ok), 'wall_s': time.perf_counter() - t0,
                'n_like': int(smp.n_like), 'n_bounds': len(smp.bounds),
                'raw_proposals': int(raw_total),
                'run': 'Sampler(prior=identity, Gaussian(30, sigma=0.1), '
                       'n_live=2000, seed=0).run(n_eff=100000, '
                       'discard_exploration=True)'}

    line = {
        'metric': 'raw_proposals_per_sec', 'value': value,
        'unit': 'proposals/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': CONFIGS[CONFIG]['scaling'],
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(n),
        'notes': {'emulator_arith': args.mlp,
                  'parallelism': 'proposal batch sharded over {} GPU(s), one '
                                 'asynchronous, double-buffered all-gather '
                                 'of 12 words per step'.format(world)},
        'roofline': roofline,
        'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': 'proposals/s',
                'h2d_bytes_per_step': bytes_io['h2d'],
                'd2h_bytes_per_step': bytes_io['d2h'],
                'ms_per_step': e2e_ms / args.steps,
                'frac_of_device': e2e_value / value,
                'api': 'nb200_session_submit/_wait_index (C ABI, host '
                       'buffers): per in-shell proposal the host receives '
                       '(global index u64, log_l f64); rows stay on the '
                       'device, nb200_session_materialize regenerates them '
                       'bit-identically on demand; 2 batches in flight, host '
                       'wall clock',
                'rows_mode': {
                    'value': world * n * rows_steps / (rows_ms * 1e-3),
                    'ms_per_step': rows_ms / rows_steps,
                    'steps': rows_steps,
                    'd2h_bytes_per_step': rows_io['d2h'],
                    'api': 'nb200_session_submit/_wait: in-shell ROWS cross '
                           'PCIe (round-1 form)'}},
        'gpu_launches': launches,
        'later_bounds': later_line,
        'delta_log_z': None if logz is None else logz['delta_log_z'],
        'log_z_run': logz,
        'clocks': clocks.summary() if clocks else None,
        'result': result,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--config', type=int, default=2, choices=sorted(CONFIGS),
                    help='2: headline, weak scaling (default); 5: 100-D, '
                         'fixed global batch, strong scaling')
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--mlp', default='f16', choices=['f64', 'tf32', 'f16'],
                    help='emulator arithmetic: tcgen05 kind::f16 (default), '
                         'kind::tf32, or the fp64 parity kernels')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-later', action='store_true',
                    help='skip the cfg2 + 47 later bounds workload')
    ap.add_argument('--no-logz', action='store_true',
                    help='skip the end-to-end config-2 run (delta_log_z)')
    args = ap.parse_args()
    select_config(args.config)
    if args.steps is None:
        args.steps = CONFIGS[CONFIG]['steps']
    if args.batch is None:
        args.batch = CONFIGS[CONFIG]['batch']
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
