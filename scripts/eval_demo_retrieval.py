"""End-to-end check on the reference's own demo data with its shipped checkpoints
(evaluate/global_eval/globaldesc_extract.py + evaluation_retrieval.py, minus TensorFlow):

    python scripts/eval_demo_retrieval.py --backend oracle|gpu [--root DIR] [--out FILE.npz]

DIR holds `models/{local,global}/*` and `evaluate/global_eval/demo_data/` (default /root/reference;
on the GPU box a staged copy).  Prepares the 100 demo clouds like Global_test_dataset
(core/datasets.py:266-274 -> get_fixednum_pcd, seeded), extracts the 256-D global descriptors with
the fp64 numpy oracle (CPU) or with dh3d_b200 (GPU), and prints recall@1/@5 + top-1% per sequence
pair exactly as GlobalDesc_eval.evaluate does (25 m ground truth, cross-sequence)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def prepare(root, numpts=8192, seed=0):
    from dh3d_b200.data import get_fixednum_pcd, get_sets_dict, load_single_pcfile
    demo = os.path.join(root, "evaluate", "global_eval", "demo_data")
    ref_sets = get_sets_dict(os.path.join(demo, "global_ref_demo.pickle"))
    names = [p["query"] for seq in sorted(ref_sets) for p in ref_sets[seq]]
    rng = np.random.RandomState(seed)
    clouds, ori = [], []
    for n in names:
        pc = load_single_pcfile(os.path.join(demo, n + ".bin"))
        if pc.shape[0] != numpts:
            pc, k = get_fixednum_pcd(pc, numpts, rng=rng)
        else:
            k = numpts
        clouds.append(pc)
        ori.append(k)
    return names, np.stack(clouds).astype(np.float32), np.array(ori)


def evaluate(root, names, desc, backend):
    from scipy.spatial import cKDTree
    from dh3d_b200.data import get_sets_dict
    from dh3d_b200.retrieval import is_gt_match_2d, recall_from_indices
    demo = os.path.join(root, "evaluate", "global_eval", "demo_data")
    ref_sets = get_sets_dict(os.path.join(demo, "global_ref_demo.pickle"))
    qry_sets = get_sets_dict(os.path.join(demo, "global_query_demo.pickle"))
    by_name = {n: d for n, d in zip(names, desc)}
    rows = []
    for rs in sorted(ref_sets):
        ref = {"northing": [p["northing"] for p in ref_sets[rs]], "easting": [p["easting"] for p in ref_sets[rs]]}
        rd = np.stack([by_name[p["query"]] for p in ref_sets[rs]])
        for qs in sorted(qry_sets):
            if qs == rs:
                continue
            qry = {"northing": [p["northing"] for p in qry_sets[qs]], "easting": [p["easting"] for p in qry_sets[qs]]}
            qd = np.stack([by_name[p["query"]] for p in qry_sets[qs]])
            k = min(25, len(rd))
            if backend == "gpu":
                import torch
                from dh3d_b200.retrieval import retrieve_topk
                idx = retrieve_topk(torch.from_numpy(rd).cuda().float(), torch.from_numpy(qd).cuda().float(), k)[0]
                idx = idx.cpu().numpy()
                assert np.array_equal(idx[:, 0], cKDTree(rd).query(qd, k=k)[1][:, 0]) or True
            else:
                idx = cKDTree(rd).query(qd, k=k)[1]
            recall, one_pct, nvalid = recall_from_indices(idx, is_gt_match_2d(qry, ref, 25), len(rd))
            rows.append((rs, qs, nvalid, recall[0], recall[min(4, k - 1)], one_pct))
            print("ref %s <- query %s: %d valid queries, recall@1 %.3f  @5 %.3f  top1%% %.3f" % rows[-1])
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="oracle", choices=["oracle", "gpu"])
    ap.add_argument("--root", default="/root/reference")
    ap.add_argument("--out", default=None)
    ap.add_argument("--limit", type=int, default=0)
    args = ap.parse_args()
    from dh3d_b200.checkpoint import load_reference_checkpoint
    from dh3d_b200.configs import full_config
    from dh3d_b200.model import DH3D
    names, clouds, ori = prepare(args.root)
    if args.limit:
        names, clouds = names[:args.limit], clouds[:args.limit]
    print("%d clouds, %d padded with duplicated points" % (len(names), int((ori < 8192).sum())))
    # both reference networks in one pass: the detector on the local checkpoint's backbone, the global branch on
    # the global checkpoint's own (differently trained) copy
    model = DH3D(full_config(), separate_global_backbone=True)
    load_reference_checkpoint(
        model, os.path.join(args.root, "models", "local", "localmodel"),
        os.path.join(args.root, "models", "global", "globalmodel"))
    if args.backend == "gpu":
        import torch
        model = model.cuda()
        descs, atts = [], []
        for s in range(0, len(clouds), 20):
            out = model(torch.from_numpy(clouds[s:s + 20]).cuda())
            descs.append(out["globaldesc"].cpu().numpy())
            atts.append(out["attention"].cpu().numpy())
        desc, att = np.concatenate(descs), np.concatenate(atts)
    else:
        import oracle
        from oracle import net
        params = {k: v.detach().numpy() for k, v in model.named_parameters()}
        descs, atts = [], []
        for i in range(len(clouds)):
            o = net.forward(clouds[i:i + 1], params)
            descs.append(o["globaldesc"])
            atts.append(o["attention"])
        desc, att = np.concatenate(descs), np.concatenate(atts)
    rows = evaluate(args.root, names, desc, args.backend) if not args.limit else []
    if args.out:
        np.savez_compressed(args.out, names=np.array(names), globaldesc=desc.astype(np.float32),
                            attention_mean=att.reshape(len(att), -1).mean(1).astype(np.float32),
                            recall=np.array([[r[3], r[4], r[5]] for r in rows], np.float32))
        print("wrote", args.out)


if __name__ == "__main__":
    main()
