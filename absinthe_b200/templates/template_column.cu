{#- Column kernel of one fused group (included by template_tiling.cu with `g` in scope; see absinthe_b200/column.py).
    Replaces the fused tile loop of the reference (ref template_tiling.cpp:301-336) for groups with k-dependences:
    register-carried streaming along k, one warp per strip of {{g.USEFUL}} columns x {{g.RJ}} rows, no CTA barriers. -#}
// column kernel: work item = {{g.B[0]}}x{{g.B[1]}} columns over the k-range of a tile; {{g.WX}}x{{g.WY}} warps, a thread owns one
// column and {{g.RJ}} row(s) and walks k.  Leads (planes ahead of the output plane):{% for n in g.NAMES %} {{n}}+{{g.LEADS[n]}}{% endfor %}.
// Ring: {{g.NS}} slots of {{g.KC}} plane(s) of every input ({{g.SLOT_BYTES}} B per slot), warp {{g.WARPS}} is the TMA producer.
{%- if g.CLUSTER %}
// Thread-block clusters of {{g.CLUSTER}} CTAs ({{g.CX}}x{{g.CY}} neighbouring tiles): a producer issues ring slot e only after every
// peer has issued slot e - {{g.SLACK}} (tick[]: one remote mbarrier arrive per peer and slot), so the CTAs of a cluster
// request the halo cells they share within the L2's reach of each other and only one request goes to HBM.
{%- endif %}
extern "C" __global__ void {% if g.CLUSTER %}__cluster_dims__({{g.CLUSTER}}, 1, 1) {% endif %}{% if g.MAXREG %}__maxnreg__({{g.MAXREG}}){% else %}__launch_bounds__({{g.LAUNCH_THREADS}}, 1){% endif %}
{{g.KERNEL}}(const __grid_constant__ Params{{g.ID}} prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const smem = reinterpret_cast<double*>(smem_raw) + {{g.FRONT}};
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
    const int lane = tid & 31;
{%- if g.CLUSTER %}
    const int clusters = (int)gridDim.x / {{g.CLUSTER}}, my_cluster = (int)blockIdx.x / {{g.CLUSTER}};
    const int my_work = my_cluster < {{g.SN_TOTAL}} ? ({{g.SN_TOTAL}} - my_cluster + clusters - 1) / clusters : 0;   // super-tiles
    uint64_t* const tick = reinterpret_cast<uint64_t*>(smem_raw + {{g.TICK_OFFSET}});
{%- else %}
    const int my_tiles = (prm.tile_count - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int my_work = my_tiles > 0 ? my_tiles * {{g.NSUB_TOTAL}} : 0;
{%- endif %}
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + {{g.BARRIER_OFFSET}});
    uint64_t* const empty = full + {{g.NS}};
    if (tid == 0) {
        for (int s = 0; s < {{g.NS}}; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], {{g.WARPS}});
        }
{%- if g.CLUSTER %}
        for (int s = 0; s < {{g.TICKS}}; ++s) mbar_init(&tick[s], {{g.CLUSTER - 1}});
{%- endif %}
        mbar_fence_init();
    }
{%- if g.CLUSTER %}
    cluster_sync_all();       // no peer may arrive on a barrier that is not initialised yet
{%- else %}
    __syncthreads();
{%- endif %}
    if (warp >= {{g.WARPS}}) {
        if (lane == 0) {
{%- if g.WAVE_SYNC %}
            // The producers of all CTAs start their n-th work item together (a grid-wide arrive/wait per item, on a
            // counter line of its own; the consumers keep draining the ring meanwhile).  Work items n of the grid are
            // neighbouring tiles: started together they stay within a few planes of each other over one column, so
            // the halo cells two tiles share are requested within the L2's reach of each other and only one of the
            // requests goes to HBM.  Left to drift (persistent CTAs, no synchronisation) none of them did (ncu:
            // dram reads = requested box bytes, 1.55x the fields).  Waiting is bounded; the last CTA to leave resets.
            int* const sync = prm.progress;
            const int lines = (int)gridDim.x * {{g.SYNC_LINES}} - 1;        // last line: departures
            const bool synced = prm.tile_step == 1 && gridDim.x > 1;
{%- endif %}
            int chunk = 0;
{%- if g.CLUSTER %}
            const int rank = (int)blockIdx.x % {{g.CLUSTER}};
            int e = 0;                                      // ring slots issued by every producer of the cluster, dead items included
            for (int n = 0; n < my_work; ++n) {
                const WorkItem w = locate{{g.ID}}(n, 1);
                const int kb = w.ok + w.bk;
                for (int c = 0; c < {{g.ITEM_CHUNKS}}; ++c, ++e) {
                    const int slot = chunk % {{g.NS}};
                    if (w.live && chunk >= {{g.NS}}) mbar_wait(&empty[slot], ((chunk / {{g.NS}}) - 1) & 1);
                    if (e >= {{g.SLACK}}) mbar_wait(&tick[(e - {{g.SLACK}}) % {{g.TICKS}}], ((e - {{g.SLACK}}) / {{g.TICKS}}) & 1);
                    if (w.live) {
                        double* const dst = smem + slot * {{g.SLOT_DOUBLES}};
                        mbar_expect_tx(&full[slot], {{g.SLOT_BYTES}});
{%- for f in g.INPUTS %}
                        tma_load_box(dst + {{f.OFFSET}}, &prm.map[{{f.MAP}}], &full[slot], (HX + w.oi + ({{f.XLO}})) & ~1, HY + w.oj + ({{f.BMIN}}), HZ + kb + ({{g.KMIN + f.QTOP}}) + c * {{g.KC}});
{%- endfor %}
                        ++chunk;
                    }
#pragma unroll
                    for (int peer = 0; peer < {{g.CLUSTER}}; ++peer)
                        if (peer != rank) mbar_arrive_remote(&tick[e % {{g.TICKS}}], peer);
                }
            }
            // a CTA's shared memory must outlive the last arrive of its peers: wait for the ticks still outstanding
            for (int d = e > {{g.SLACK}} ? e - {{g.SLACK}} : 0; d < e; ++d) mbar_wait(&tick[d % {{g.TICKS}}], (d / {{g.TICKS}}) & 1);
{%- else %}
            for (int n = 0; n < my_work; ++n) {
{%- if g.WAVE_SYNC %}
                if (synced && n > 0 && n < lines) {
                    const int want = min((int)gridDim.x, prm.tile_count - n * (int)gridDim.x);
                    atomicAdd(sync + n * 32, 1);
                    for (int spin = 0; spin < 50000 && peek(sync + n * 32) < want; ++spin) __nanosleep(200);
                }
{%- endif %}
                const WorkItem w = locate{{g.ID}}(n, prm.tile_step);
                if (!w.live) continue;
                const int kb = w.ok + w.bk, ke = w.ok + w.ek;
                const int nchunks = ((ke - kb) - ({{g.KMIN}}) + {{g.KC - 1}}) / {{g.KC}};
                for (int c = 0; c < nchunks; ++c, ++chunk) {
                    const int slot = chunk % {{g.NS}};
                    if (chunk >= {{g.NS}}) mbar_wait(&empty[slot], ((chunk / {{g.NS}}) - 1) & 1);
                    double* const dst = smem + slot * {{g.SLOT_DOUBLES}};
                    mbar_expect_tx(&full[slot], {{g.SLOT_BYTES}});
{%- for f in g.INPUTS %}
                    tma_load_box(dst + {{f.OFFSET}}, &prm.map[{{f.MAP}}], &full[slot], (HX + w.oi + ({{f.XLO}})) & ~1, HY + w.oj + ({{f.BMIN}}), HZ + kb + ({{g.KMIN + f.QTOP}}) + c * {{g.KC}});
{%- endfor %}
                }
            }
{%- endif %}
{%- if g.WAVE_SYNC %}
            if (synced && atomicAdd(sync + lines * 32, 1) == (int)gridDim.x - 1) {
                for (int n = 0; n <= lines; ++n) post(sync + n * 32, 0);      // clean counters for the next launch
            }
{%- endif %}
        }
        return;
    }
    const int sx = warp % {{g.WX}}, sy = warp / {{g.WX}};
    const int col = sx * {{g.USEFUL}} + lane - {{g.FIRST}};      // column relative to the work item's origin
    const int row0 = sy * {{g.RJ}};
    const bool lane_owns = lane >= {{g.FIRST}} && lane <= {{g.LAST}};
    int chunk = 0;
    for (int n = 0; n < my_work; ++n) {
        const WorkItem w = locate{{g.ID}}(n, prm.tile_step);
        if (!w.live) continue;
        const int kb = w.ok + w.bk, ke = w.ok + w.ek;        // owned planes, interior coordinates
        const int steps = (ke - kb) - ({{g.KMIN}});
        const int nchunks = (steps + {{g.KC - 1}}) / {{g.KC}};
        const bool okc = lane_owns && col >= w.bi && col < w.ei;
{%- for b in range(g.RJ) %}
        const bool ok_{{b}} = okc && row0 + {{b}} >= w.bj && row0 + {{b}} < w.ej;
{%- endfor %}
{%- for f in g.INPUTS %}
        const int off_{{f.NAME}} = {{f.OFFSET}} + row0 * {{f.ROW}} + sx * {{g.USEFUL}} + lane + ((HX + w.oi + ({{f.XLO}})) & 1);
{%- endfor %}
        const long origin = (HX + w.oi + col) + (long)(HY + w.oj + row0) * ROW + (long)HZ * PLANE;
{%- for o in g.OUT_LEADS %}
        double* const out_{{o.NAME}} = prm.out[{{o.INDEX}}] + origin;
{%- endfor %}
{%- for c in g.CARRIED %}
        double {{c}} = 0.0;
{%- endfor %}

        // {{g.UNROLL}} ring slot(s) per trip of the loop, fully unrolled: the carried registers are renamed from step to
        // step instead of being copied
        for (int c0 = 0; c0 < nchunks; c0 += {{g.UNROLL}}) {
#pragma unroll
            for (int m = 0; m < {{g.UNROLL}}; ++m) {
                const int c = c0 + m;
                if (c < nchunks) {                               // warp-uniform
                    const int slot = chunk % {{g.NS}};
                    mbar_wait(&full[slot], (chunk / {{g.NS}}) & 1);
                    const double* const base = smem + slot * {{g.SLOT_DOUBLES}};
#pragma unroll
                    for (int r = 0; r < {{g.KC}}; ++r) {
                        const int t = c * {{g.KC}} + r;
                        if (t < steps) {                         // warp-uniform
                            const int kk = kb + ({{g.KMIN}}) + t;    // output plane of this step
{%- for f in g.INPUTS %}
                            const double* const pl_{{f.NAME}} = base + off_{{f.NAME}} + r * {{f.PLANE}};
{%- endfor %}
{%- for o in g.OUT_LEADS %}
                            const int p_{{o.NAME}} = kk + ({{o.LEAD}});
                            const bool pk_{{o.NAME}} = p_{{o.NAME}} >= kb && p_{{o.NAME}} < ke;
                            double* const o_{{o.NAME}} = out_{{o.NAME}} + (long)p_{{o.NAME}} * PLANE;
{%- endfor %}
{%- for line in g.BODY %}
                            {{line}}
{%- endfor %}
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[slot]);    // this warp is done with the slot
                    ++chunk;
                }
            }
        }
    }
}
