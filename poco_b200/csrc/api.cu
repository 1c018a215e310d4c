// poco_b200 -- C-ABI entry points that are not tied to one kernel file: error reporting, device
// check, op dispatch and the plan (static layer schedule) executor.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"

namespace poco {

std::atomic<int64_t> g_launches{0};
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

}  // namespace poco

using namespace poco;

constexpr int kMaxLanes = 8;

struct poco_plan {
    std::vector<poco_op> ops;
    int64_t flops = 0;
    cudaStream_t side[kMaxLanes] = {};         // lanes 1..7 (lane 0 is the caller's stream)
    std::vector<cudaEvent_t> events;           // one per fork, one per (join, lane)
    ~poco_plan() {
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        for (cudaStream_t s : side)
            if (s) cudaStreamDestroy(s);
    }
};

extern "C" int poco_version(void) { return 100; }

extern "C" const char* poco_last_error(void) { return g_error.c_str(); }

extern "C" int64_t poco_kernel_launches(void) { return g_launches.load(); }

extern "C" int poco_device_check(int device) {
    cudaDeviceProp prop;
    POCO_CUDA(cudaGetDeviceProperties(&prop, device));
    POCO_CHECK(prop.major == 10, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                     ", poco_b200 kernels are sm_100a only");
    return 0;
}

extern "C" int poco_conv_run(const poco_conv* d, void* stream) {
    if (check_act(d->in, "in")) return 1;
    if (d->s2d_only && d->out.data == nullptr) {        // geometry-only: nothing is written through `out`
        POCO_CHECK(d->out.C > 0 && d->out.C % 8 == 0 && d->out.N > 0 && d->out.H > 0 && d->out.W > 0 &&
                       d->out.plane_stride >= int64_t(d->out.N) * (d->out.H + 2) * (d->out.W + 2), "out: bad geometry");
    } else if (check_act(d->out, "out")) {
        return 1;
    }
    POCO_CHECK(d->weight && d->bias, "null weight / bias");
    POCO_CHECK(d->kh >= 1 && d->kw >= 1 && d->pad >= 0, "bad kernel geometry");
    POCO_CHECK(!d->residual || d->res_plane_stride >= int64_t(d->out.N) * (d->out.H + 2) * (d->out.W + 2),
               "residual plane stride too small");
    POCO_CHECK(d->impl == 0 || (d->wfmt == 0 && !d->in_s2d && d->out_s2d.data == nullptr),
               "the debug kernel reads the standard weight layout and knows no space-to-depth plumbing");
    if (d->out_s2d.data != nullptr && check_act(d->out_s2d, "out_s2d")) return 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return d->impl == 1 ? conv_ref_launch(d, s) : conv_tc_launch(d, s);
}

extern "C" int64_t poco_conv_chain_flag_count(const poco_conv_chain* d) {
    if (!d || d->n_seg < 1) return 0;
    const poco_act& o = d->seg[0].out;
    const int64_t tiles = (int64_t(o.N) * (o.H + 2) * (o.W + 2) + 127) / 128;
    return int64_t(d->n_seg - 1) * tiles;
}

extern "C" int poco_conv_chain_run(const poco_conv_chain* d, void* stream) {
    POCO_CHECK(d->n_seg >= 1 && d->n_seg <= POCO_MAX_CHAIN, "bad chain length");
    for (int i = 0; i < d->n_seg; ++i) {
        const poco_conv& c = d->seg[i];
        if (check_act(c.in, "in") || check_act(c.out, "out")) return 1;
        POCO_CHECK(c.weight && c.bias, "null weight / bias");
        POCO_CHECK(c.impl == 0, "a chain runs on the tcgen05 kernel only");
        POCO_CHECK(c.max_ctas == d->seg[0].max_ctas, "chain: segments must share one CTA budget");
        POCO_CHECK(!c.residual || c.res_plane_stride >= int64_t(c.out.N) * (c.out.H + 2) * (c.out.W + 2),
                   "residual plane stride too small");
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (d->n_seg > 1) {
        POCO_CHECK(d->flags != nullptr, "chain: null flags");
        POCO_CUDA(cudaMemsetAsync(d->flags, 0, size_t(poco_conv_chain_flag_count(d)) * sizeof(int32_t), s));
    }
    return conv_tc_launch_chain(d->seg, d->n_seg, d->flags, s);
}

extern "C" int poco_run_op(const poco_op* op, void* stream) {
    switch (op->kind) {
        case POCO_OP_PACK_IMAGE: return poco_pack_image_run(&op->u.pack_image, stream);
        case POCO_OP_CONV: return poco_conv_run(&op->u.conv, stream);
        case POCO_OP_CONV_CHAIN: return poco_conv_chain_run(&op->u.conv_chain, stream);
        case POCO_OP_FUSE_SUM: return poco_fuse_sum_run(&op->u.fuse_sum, stream);
        case POCO_OP_UPSAMPLE2X: return poco_upsample2x_run(&op->u.upsample2x, stream);
        case POCO_OP_MAXPOOL: return poco_maxpool_run(&op->u.maxpool, stream);
        case POCO_OP_AVGPOOL: return poco_avgpool_run(&op->u.avgpool, stream);
        case POCO_OP_UNPACK: return poco_unpack_run(&op->u.unpack, stream);
        case POCO_OP_LINEAR: return poco_linear_run(&op->u.linear, stream);
        case POCO_OP_COPY2D: return poco_copy2d_run(&op->u.copy2d, stream);
        case POCO_OP_ROT6D: return poco_rot6d_run(&op->u.rot6d, stream);
        case POCO_OP_PARE_HEAD: return poco_pare_head_run(&op->u.pare_head, stream);
        case POCO_OP_REALNVP: return poco_realnvp_run(&op->u.realnvp, stream);
        case POCO_OP_CROP: return poco_crop_run(&op->u.crop, stream);
        case POCO_OP_UNCERT_POST: return poco_uncert_post_run(&op->u.uncert_post, stream);
        case POCO_OP_SMPL: return poco_smpl_run(&op->u.smpl, stream);
        case POCO_OP_BASIC_BLOCK: return poco_basic_block_run(&op->u.basic_block, stream);
        case POCO_OP_BOTTLENECK_TAIL: return poco_bottleneck_tail_run(&op->u.bottleneck_tail, stream);
        case POCO_OP_BRANCH: return poco_branch_run(&op->u.branch, stream);
        default: break;
    }
    set_error("poco_run_op: unknown op kind " + std::to_string(op->kind));
    return 1;
}

extern "C" int poco_plan_create(const poco_op* ops, int32_t n_ops, poco_plan** out) {
    POCO_CHECK(ops != nullptr && n_ops > 0 && out != nullptr, "bad arguments");
    poco_plan* p = new poco_plan();
    p->ops.assign(ops, ops + n_ops);
    for (const poco_op& op : p->ops) {
        if (op.kind == POCO_OP_CONV) p->flops += conv_flops(&op.u.conv);
        if (op.kind == POCO_OP_CONV_CHAIN)
            for (int k = 0; k < op.u.conv_chain.n_seg; ++k) p->flops += conv_flops(&op.u.conv_chain.seg[k]);
        if (op.kind == POCO_OP_BASIC_BLOCK) {
            const poco_act& a = op.u.basic_block.out;
            p->flops += 2 * (2ll * a.N * a.H * a.W * a.C * a.C * 9);
        }
        if (op.kind == POCO_OP_BOTTLENECK_TAIL) {
            const poco_bottleneck_tail& t = op.u.bottleneck_tail;
            p->flops += 2ll * t.out.N * t.out.H * t.out.W * (int64_t(t.in.C) * t.in.C * 9 + int64_t(t.in.C) * t.out.C);
        }
        if (op.kind == POCO_OP_LINEAR) p->flops += 2ll * op.u.linear.M * op.u.linear.I * op.u.linear.O;
        if (op.kind == POCO_OP_BRANCH) {
            const poco_act& a = op.u.branch.out;
            p->flops += 2 * op.u.branch.n_blocks * (2ll * a.N * a.H * a.W * a.C * a.C * 9);
        }
        if (op.kind < POCO_OP_PACK_IMAGE || op.kind > POCO_OP_BRANCH || op.lane < 0 || op.lane >= kMaxLanes) {
            delete p;
            set_error("poco_plan_create: unknown op kind " + std::to_string(op.kind));
            return 1;
        }
    }
    int max_lane = 0, n_events = 0;
    for (const poco_op& op : p->ops) {
        max_lane = std::max(max_lane, int(op.lane));
        if (op.kind == POCO_OP_FORK) { max_lane = std::max(max_lane, op.u.sync.n_lanes - 1); n_events += 1; }
        if (op.kind == POCO_OP_JOIN) n_events += op.u.sync.n_lanes;
    }
    if (max_lane >= kMaxLanes) {
        delete p;
        set_error("poco_plan_create: too many lanes");
        return 1;
    }
    // CUDA stream priorities of lanes 1, 2, ... (lower = served first when an SM frees up; lane 0 is the caller's stream).
    // The lanes of an HR module oversubscribe the SMs, so the block scheduler decides which lane's queued CTAs run next.
    // The low-resolution branches are dependent chains of short launches and end the module, the high-resolution lane 0
    // has the long kernels that fill whatever is free: serving the side lanes first measured 10.78 -> 10.59 ms per step
    // (cliff_w32, batch 256; stream capture records the priority in the graph's kernel nodes).
    // POCO_B200_LANE_PRIO="p1,p2,..." overrides the default -1,-2,-3,...; "off" = plain streams.
    int prio[kMaxLanes] = {};
    bool use_prio = false;
    const char* env = getenv("POCO_B200_LANE_PRIO");
    if (env == nullptr) env = "-1,-2,-3,-4,-5,-6,-7";
    if (strcmp(env, "off") != 0) {
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) {
            use_prio = true;
            const char* q = env;
            for (int k = 1; k < kMaxLanes && *q; ++k) {
                char* end = nullptr;
                const long v = strtol(q, &end, 10);
                if (end == q) break;
                prio[k] = int(std::min<long>(least, std::max<long>(greatest, v)));
                q = *end == ',' ? end + 1 : end;
            }
        } else {
            cudaGetLastError();
        }
    }
    for (int k = 1; k <= max_lane; ++k)
        if ((use_prio ? cudaStreamCreateWithPriority(&p->side[k], cudaStreamNonBlocking, prio[k])
                      : cudaStreamCreateWithFlags(&p->side[k], cudaStreamNonBlocking)) != cudaSuccess) {
            // (no device in host-logic tests: lanes then run on the caller's stream)
            p->side[k] = nullptr;
            cudaGetLastError();
        }
    p->events.resize(n_events, nullptr);
    for (cudaEvent_t& e : p->events)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            e = nullptr;
            cudaGetLastError();
        }
    *out = p;
    return 0;
}

extern "C" int poco_plan_run(poco_plan* plan, void* stream) {
    POCO_CHECK(plan != nullptr, "null plan");
    cudaStream_t main_s = static_cast<cudaStream_t>(stream);
    size_t ev = 0;
    for (size_t i = 0; i < plan->ops.size(); ++i) {
        const poco_op& op = plan->ops[i];
        if (op.kind == POCO_OP_FORK) {
            cudaEvent_t e = plan->events[ev++];
            if (e) {
                POCO_CUDA(cudaEventRecord(e, main_s));
                for (int k = 1; k < op.u.sync.n_lanes; ++k)
                    if (plan->side[k]) POCO_CUDA(cudaStreamWaitEvent(plan->side[k], e, 0));
            }
            continue;
        }
        if (op.kind == POCO_OP_JOIN) {
            for (int k = 1; k < op.u.sync.n_lanes; ++k) {
                cudaEvent_t e = plan->events[ev++];
                if (e && plan->side[k]) {
                    POCO_CUDA(cudaEventRecord(e, plan->side[k]));
                    POCO_CUDA(cudaStreamWaitEvent(main_s, e, 0));
                }
            }
            ++ev;       // (a join reserves n_lanes events; slot 0 is unused)
            continue;
        }
        cudaStream_t s = (op.lane > 0 && plan->side[op.lane]) ? plan->side[op.lane] : main_s;
        const int rc = poco_run_op(&op, s);
        if (rc != 0) {
            set_error("op " + std::to_string(i) + " (kind " + std::to_string(op.kind) + "): " + g_error);
            return rc;
        }
    }
    return 0;
}

extern "C" int32_t poco_plan_num_ops(const poco_plan* plan) { return plan ? int32_t(plan->ops.size()) : 0; }
extern "C" int64_t poco_plan_flops(const poco_plan* plan) { return plan ? plan->flops : 0; }
extern "C" void poco_plan_destroy(poco_plan* plan) { delete plan; }
