"""Per-stream time line of one eagerly executed step (no nsys in this image): which chain is the critical path, where streams idle.

    python tools/timeline.py [--workload image|video] [--batch N] > gpurun_out/timeline.txt
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/timeline.py   (data parallel: rank 0's time line;
        stream 4 = the bulk communicator's stream with the gradient-bucket all-reduces)
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="image")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--bn-local", action="store_true")
    args = ap.parse_args()
    import video_filler_b200.tensor as T
    from video_filler_b200 import models, synth, train, util
    world, rank, local_rank, dist = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), None
    T.state(local_rank)
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idbuf = np.zeros(128, np.uint8)
        if rank == 0:
            T.api().cenn_dist_unique_id(idbuf.ctypes.data_as(C.c_void_p))
        idt = torch.from_numpy(idbuf).cuda()
        dist.broadcast(idt, src=0)
        idbuf = idt.cpu().numpy()
        T.api().cenn_dist_init(T.state(), idbuf.ctypes.data_as(C.c_void_p), world, rank)
    video = args.workload == "video"
    B = args.batch or (64 if video else 256)
    opt = models.default_opt("video" if video else "image", batchSize=B)
    if video:
        opt["wtgdl"] = 0.5
    trn = train.FusedTrainer(opt, precision="bf16", world_size=world, rank=rank, bn_local=1 if args.bn_local else 0)
    rng = np.random.default_rng(1234)
    trn.set_params(0, util.params_flat(util.weights_init(util.describe_netG(opt), rng)))
    trn.set_params(1, util.params_flat(util.weights_init(util.describe_netD(opt), rng)))
    if video:
        ctx, full, mask = synth.video_batch(B, 12, 128, opt["maskValue"], rng)
        da, db = T.CudaTensor.from_numpy(ctx), T.CudaTensor.from_numpy(full)
        api, st = T.api(), T.state()
        mask = np.ascontiguousarray(mask, np.uint8)
        p = C.c_void_p(); api.cenn_malloc(st, mask.nbytes, C.byref(p)); api.cenn_copy_h2d(st, p, mask.ctypes.data_as(C.c_void_p), mask.nbytes)
        mptr = p.value
    else:
        ctx, center = synth.image_batch(B, 128, 4, rng)
        da, db, mptr = T.CudaTensor.from_numpy(ctx), T.CudaTensor.from_numpy(center), None
    for _ in range(3):
        trn.step_device(da.ptr, db.ptr, mptr)
    tl = trn.timeline(da.ptr, db.ptr, mptr)
    if rank != 0:
        trn.close(); T.api().cenn_dist_shutdown(T.state()); dist.barrier(); dist.destroy_process_group()
        return
    end = max(x[3] for x in tl)
    print("# eager step, %s B=%d: %.3f ms from first op to last completion" % (args.workload, B, end))
    print("# %-4s %-18s %2s %9s %9s %8s" % ("idx", "op", "st", "start", "end", "dur"))
    busy = {}
    for i, (nm, sid, a, b) in enumerate(tl):
        print("%5d %-18s %2d %9.4f %9.4f %8.4f" % (i, nm, sid, a, b, b - a))
        busy[sid] = busy.get(sid, 0.0) + (b - a)
    print("# busy ms per stream:", {k: round(v, 3) for k, v in sorted(busy.items())})
    trn.close()
    if world > 1:
        T.api().cenn_dist_shutdown(T.state()); dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
