"""Training / evaluation driver with the reference's flags and loop (reference ``kgat.py``).

    python -m dgl_kgat_b200.kgat --data_name amazon-book            # datasets/<name>/data/*.pd
    python -m dgl_kgat_b200.kgat --synthetic tiny --max_epoch 2     # no files needed

Every flag of kgat.py:12-41 is kept with the same name and default.  Per epoch, as in
kgat.py:114-196: (1) TransR on the KG in batches of ``--batch_size_kg``; (2) attention refresh
under ``no_grad`` stored as ``train_g.edata['w']``; (3) CF phase: full-graph ``gnn`` forward +
BPR loss + backward per batch of ``--batch_size``; (4) every ``--evaluate_every`` epochs
Recall@20 / NDCG@20 on the validation users (train graph) and the test users (test graph),
logged to ``log/<cfg>/valid{id}.csv`` and ``test{id}.csv``.  Additions (all off by default):
``--synthetic`` builds the inputs from :mod:`synth`; ``--max_kg_iters`` / ``--max_cf_iters``
cap the phases for smoke runs; batches come from the device-resident sampler unless
``--host_sampler`` is given; the KG and CF iterations run as CUDA graphs (``graphed.py``) unless
``--eager``; Adam is one library launch per step (``optim.FusedAdam``); ``--save_every`` / ``--resume`` write and read
checkpoints (the reference has none, kgat.py:190).

Multi-GPU (the reference is single-GPU): launched as
``python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P -m dgl_kgat_b200.kgat ...``
the same loop runs with the attention refresh, the CF phase and the evaluation on the destination-partitioned graph
(``dist.DistKGAT``); parameters stay bit-identical across ranks, rank 0 writes the logs and checkpoints.
"""
from __future__ import annotations

import argparse
import logging
import os
from time import time

import torch as th
import torch.optim as optim

from . import metric
from .dataset import DataLoader
from .models import Model
from .optim import FusedAdam
from .utils import MetricLogger, creat_log_id, logging_config


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="KGAT on the B200-native hot path")
    p.add_argument('--gpu', type=int, default=0, help='use GPU')
    p.add_argument('--seed', type=int, default=1234, help='the random seed')
    p.add_argument('--data_name', nargs='?', default='yelp2018', help='Choose a dataset from {yelp2018, amazon-book}')
    p.add_argument('--entity_embed_dim', type=int, default=64, help='KG entity Embedding size.')
    p.add_argument('--relation_embed_dim', type=int, default=64, help='KG relation Embedding size.')
    p.add_argument('--gnn_model', type=str, default="kgat", help='the gnn models')
    p.add_argument('--gnn_num_layer', type=int, default=3, help='the number of layers')
    p.add_argument('--gnn_hidden_size', type=int, default=64, help='Output sizes of every layer')
    p.add_argument('--dropout_rate', type=float, default=0.1, help='dropout probability of every layer output')
    p.add_argument('--use_attention', type=bool, default=True, help='Whether to use attention to update adj')
    p.add_argument('--regs', type=float, default=0.0001, help='Regularization for user and item embeddings.')
    p.add_argument('--max_epoch', type=int, default=5000, help='train xx iterations')
    p.add_argument("--grad_norm", type=float, default=1.0, help="norm to clip gradient to (unused, as in the reference)")
    p.add_argument('--lr', type=float, default=0.001, help='Learning rate.')
    p.add_argument('--batch_size', type=int, default=10240, help='CF batch size.')
    p.add_argument('--batch_size_kg', type=int, default=2048, help='KG batch size.')
    p.add_argument('--joint_train', type=bool, default=False, help='unused, as in the reference')
    p.add_argument('--evaluate_every', type=int, default=1, help='the evaluation duration')
    p.add_argument('--print_every', type=int, default=200, help='the print duration')
    # additions
    p.add_argument('--synthetic', type=str, default=None, help='synthetic CKG config (tiny, last-fm, yelp2018, amazon-book, ...)')
    p.add_argument('--max_kg_iters', type=int, default=-1, help='cap on KG batches per epoch (-1: all)')
    p.add_argument('--max_cf_iters', type=int, default=-1, help='cap on CF batches per epoch (-1: all)')
    p.add_argument('--log_root', type=str, default='log')
    p.add_argument('--host_sampler', action='store_true',
                   help='sample batches with numpy on the host (default: device-resident sampler)')
    p.add_argument('--eager', action='store_true',
                   help='do not capture the KG / CF iterations as CUDA graphs')
    p.add_argument('--save_every', type=int, default=0, help='write a checkpoint every N epochs (0: never)')
    p.add_argument('--resume', type=str, default=None, help='checkpoint file to resume from')
    args = p.parse_args(argv)
    name = args.data_name if args.synthetic is None else "synthetic-" + args.synthetic
    save_dir = "{}_d{}_l{}_dp{}_lr{}_bz{}_kgbz{}_att{}_seed{}".format(
        name, args.entity_embed_dim, args.gnn_num_layer, args.dropout_rate, args.lr,
        args.batch_size, args.batch_size_kg, int(args.use_attention), args.seed)
    args.save_dir = os.path.join(args.log_root, save_dir)
    # under torchrun (one process per GPU) every rank but the first logs into its own sub-folder
    args.rank, args.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if args.world > 1:
        args.gpu = int(os.environ.get("LOCAL_RANK", args.rank))
        if args.rank > 0:
            args.save_dir = os.path.join(args.save_dir, "rank%d" % args.rank)
    os.makedirs(args.save_dir, exist_ok=True)
    args.save_id = creat_log_id(args.save_dir)
    return args


def eval(model, g, train_user_dict, eval_user_dict, item_id_range, use_cuda, use_attention):
    """kgat.py:53-62: refresh attention, propagate, rank."""
    with th.no_grad():
        if use_attention:
            A_w = model.compute_attention(g)
            g.edata['w'] = A_w
        all_embedding = model.gnn(g, g.ndata['id'])
        return metric.calc_recall_ndcg(all_embedding, train_user_dict, eval_user_dict, item_id_range,
                                       K=20, use_cuda=use_cuda)


def eval_dist(dk, train_user_dict, eval_user_dict, item_id_range, use_attention):
    """eval() on the destination-partitioned graph: attention and propagation partitioned, the embedding rows
    all-gathered, the eval users sharded over the ranks, three scalars all-reduced."""
    with th.no_grad():
        if use_attention:
            dk.compute_attention()
        emb = dk.full_embedding(dk.gnn())        # dropout follows model.training, as in the reference's eval()
        return dk.calc_recall_ndcg(emb, train_user_dict, eval_user_dict, item_id_range, 20)


def train_eval(args):
    rank, world = getattr(args, "rank", 0), getattr(args, "world", 1)
    logging_config(folder=args.save_dir, name='log{:d}'.format(args.save_id), no_console=rank > 0)
    logging.info(args)
    if not (args.gpu >= 0 and th.cuda.is_available()):
        raise RuntimeError("the KGAT hot path needs a B200 GPU; there is no CPU fallback")
    th.cuda.set_device(args.gpu)
    dev = th.device("cuda", args.gpu)
    th.manual_seed(args.seed)
    if world > 1:
        # Multi-GPU (python -m torch.distributed.run ... -m dgl_kgat_b200.kgat ...): the CF phase, the attention refresh
        # and the evaluation run on the destination-partitioned graph (dist.DistKGAT: one process per GPU, exchange
        # over NVLink peer memory); parameters and optimizer state are replicated and stay bit-identical because
        # every rank draws the same batches and masks from the same seed and every gradient is complete and
        # identical on every rank.  The TransR phase is replicated.
        import torch.distributed as tdist
        if not tdist.is_initialized():
            tdist.init_process_group("nccl", device_id=dev)

    if args.synthetic is not None:
        dataset = DataLoader.from_synthetic(args.synthetic, seed=args.seed)
    else:
        dataset = DataLoader(data_name=args.data_name, seed=args.seed)
    print(dataset)
    if not args.host_sampler:
        dataset.use_device_samplers(dev)
    model = Model(use_KG=True, input_node_dim=args.entity_embed_dim, gnn_model=args.gnn_model,
                  num_gnn_layers=args.gnn_num_layer, n_hidden=args.gnn_hidden_size, dropout=args.dropout_rate,
                  n_entities=dataset.n_KG_entity, n_relations=dataset.n_KG_relation,
                  relation_dim=args.relation_embed_dim, reg_lambda_kg=args.regs, reg_lambda_gnn=args.regs).to(dev)
    logging.info(model)
    optimizer = FusedAdam(model.parameters(), lr=args.lr)     # optim.Adam's rule, one library launch per step
    start_epoch = 1
    if args.resume:
        ckpt = th.load(args.resume, map_location=dev)
        model.load_state_dict(ckpt["state_dict"])
        optimizer.load_state_dict(ckpt["optimizer"])
        start_epoch = int(ckpt["epoch"]) + 1
        logging.info("resumed from %s (epoch %d)", args.resume, ckpt["epoch"])
    valid_log = MetricLogger(['epoch', 'recall', 'ndcg', 'is_best'], ['%d', '%.5f', '%.5f', '%d'],
                             os.path.join(args.save_dir, 'valid{:d}.csv'.format(args.save_id)))
    test_log = MetricLogger(['epoch', 'recall', 'ndcg'], ['%d', '%.5f', '%.5f'],
                            os.path.join(args.save_dir, 'test{:d}.csv'.format(args.save_id)))
    best_epoch, best_recall = -1, 0.0
    test_recall = test_ndcg = float("nan")       # test metrics AT THE BEST validation epoch (kgat.py:179-198)
    if args.resume:
        best_epoch, best_recall = int(ckpt.get("best_epoch", -1)), float(ckpt.get("best_recall", 0.0))
        test_recall, test_ndcg = float(ckpt.get("test_recall", test_recall)), float(ckpt.get("test_ndcg", test_ndcg))
        if "torch_rng" in ckpt:
            th.set_rng_state(ckpt["torch_rng"].cpu())
        if "cuda_rng" in ckpt:
            th.cuda.set_rng_state(ckpt["cuda_rng"].cpu(), dev)

    train_g, test_g = dataset.train_g, dataset.test_g
    for g in (train_g, test_g):                              # kgat.py:95-109
        g.ndata['id'] = g.ndata['id'].to(dev)
        g.edata['type'] = g.edata['type'].to(dev)
    item_id_range = th.as_tensor(dataset.item_id_range, dtype=th.long, device=dev)
    as_ids = lambda a: th.as_tensor(a, dtype=th.long).to(dev, non_blocking=True)

    dk_train = dk_test = None
    if world > 1:
        from .dist import DistKGAT
        N, R = dataset.n_KG_entity, dataset.n_KG_relation
        dk_train = DistKGAT(model, N, train_g.triplets, R, dev, max_batch=args.batch_size)
        dk_test = DistKGAT(model, N, test_g.triplets, R, dev, max_batch=args.batch_size)
        logging.info("rank %d of %d owns rows [%d, %d) of the train graph (%s exchange)", rank, world,
                     dk_train.part.r0, dk_train.part.r1, dk_train.exchange)

    kg_graph = cf_graph = None
    for epoch in range(start_epoch, args.max_epoch + 1):
        # phase 1: TransR on the KG (kgat.py:116-136)
        t0, it, total = time(), 0, th.zeros((), dtype=th.float64, device=dev)
        model.train()
        for h, r, pos_t, neg_t, _ in dataset.KG_sampler(batch_size=args.batch_size_kg):
            if 0 <= args.max_kg_iters <= it:
                break
            it += 1
            if not args.eager and len(h) == args.batch_size_kg:
                if kg_graph is None:
                    from .graphed import GraphedKGStep
                    kg_graph = GraphedKGStep(model, args.batch_size_kg)
                loss = kg_graph(h, r, pos_t, neg_t)
            else:
                loss = model.transR(as_ids(h), as_ids(r), as_ids(pos_t), as_ids(neg_t))
                loss.backward()
                # drop the eager autograd graph at once: while it lives, the parameters' AccumulateGrad nodes stay
                # bound to this stream and torch refuses to capture a graphed step (the last, partial batch of an
                # epoch is always eager, and the CF step is captured right after it in epoch 1)
                loss = loss.detach()
            optimizer.step()
            optimizer.zero_grad()
            total.add_(loss.detach())               # summed on the device: no host sync per iteration
            if it % args.print_every == 0 or it == 1:
                logging.info("Epoch {:04d} Iter {:04d} | Loss {:.4f} ".format(epoch, it, total.item() / it))
        mean_loss = total.item() / max(it, 1)            # the read-back waits for the phase: the time below is real
        logging.info('Time for KGE: {:.1f}s, loss {:.4f}'.format(time() - t0, mean_loss))

        # phase 2: attention refresh (kgat.py:139-145)
        if args.use_attention:
            t0 = time()
            if dk_train is not None:
                dk_train.compute_attention()
            else:
                with th.no_grad():
                    A_w = model.compute_attention(train_g)
                train_g.edata['w'] = A_w
            th.cuda.synchronize()
            logging.info("Time for attention: {:.3f}s".format(time() - t0))

        # phase 3: CF (kgat.py:146-168)
        t0, it, total = time(), 0, th.zeros((), dtype=th.float64, device=dev)
        for user_ids, item_pos_ids, item_neg_ids, _ in dataset.CF_pair_sampler(batch_size=args.batch_size):
            if 0 <= args.max_cf_iters <= it:
                break
            it += 1
            if not args.eager and len(user_ids) == args.batch_size:
                if cf_graph is None:
                    from .graphed import GraphedCFStep, GraphedDistCFStep
                    cf_graph = GraphedCFStep(model, train_g, args.batch_size, refresh_attention=False) \
                        if dk_train is None else GraphedDistCFStep(dk_train, args.batch_size, refresh_attention=False)
                loss = cf_graph(user_ids, item_pos_ids, item_neg_ids)
            elif dk_train is not None:
                embedding = dk_train.gnn()
                loss = dk_train.get_loss(embedding, as_ids(user_ids), as_ids(item_pos_ids), as_ids(item_neg_ids))
                loss.backward()
                loss, embedding = loss.detach(), None
            else:
                embedding = model.gnn(train_g, train_g.ndata['id'])
                loss = model.get_loss(embedding, as_ids(user_ids), as_ids(item_pos_ids), as_ids(item_neg_ids))
                loss.backward()
                loss, embedding = loss.detach(), None           # see the KG phase: no eager graph may outlive its step
            optimizer.step()
            optimizer.zero_grad()
            total.add_(loss.detach())               # summed on the device: no host sync per iteration
            if it % args.print_every == 0 or it == 1:
                logging.info("Epoch {:04d} Iter {:04d} | Loss {:.4f} ".format(epoch, it, total.item() / it))
        mean_loss = total.item() / max(it, 1)
        logging.info('Time for GNN: {:.2f}s, loss {:.4f}'.format(time() - t0, mean_loss))

        if epoch % args.evaluate_every == 0:                 # kgat.py:171-196
            t0 = time()
            if dk_train is not None:
                val_recall, val_ndcg = eval_dist(dk_train, dataset.train_user_dict, dataset.valid_user_dict,
                                                 dataset.item_id_range, args.use_attention)
            else:
                val_recall, val_ndcg = eval(model, train_g, dataset.train_user_dict, dataset.valid_user_dict,
                                            item_id_range, True, args.use_attention)
            info = "Epoch{}, [{:.1f}s] val recall:{:.5f}, val ndcg:{:.5f}".format(epoch, time() - t0, val_recall, val_ndcg)
            is_best = val_recall > best_recall
            valid_log.log(epoch=epoch, recall=val_recall, ndcg=val_ndcg, is_best=int(is_best))
            t0 = time()
            # the reference evaluates the test users in both branches (kgat.py:185,193) but keeps the numbers it
            # reports at the end only when the validation recall improved (kgat.py:181-187)
            if dk_test is not None:
                cur_recall, cur_ndcg = eval_dist(dk_test, dataset.train_valid_user_dict, dataset.test_user_dict,
                                                 dataset.item_id_range, args.use_attention)
            else:
                cur_recall, cur_ndcg = eval(model, test_g, dataset.train_valid_user_dict, dataset.test_user_dict,
                                            item_id_range, True, args.use_attention)
            if is_best:
                best_recall, best_epoch = val_recall, epoch
                test_recall, test_ndcg = cur_recall, cur_ndcg
                test_log.log(epoch=epoch, recall=test_recall, ndcg=test_ndcg)
            info += "\t[{:.1f}s] test recall:{:.5f}, test ndcg:{:.5f}".format(time() - t0, cur_recall, cur_ndcg)
            logging.info(info)
        if args.save_every > 0 and epoch % args.save_every == 0 and rank == 0:
            path = os.path.join(args.save_dir, 'ckpt{:d}.pt'.format(args.save_id))
            th.save({'state_dict': model.state_dict(), 'optimizer': optimizer.state_dict(), 'epoch': epoch,
                     'best_epoch': best_epoch, 'best_recall': best_recall, 'test_recall': test_recall,
                     'test_ndcg': test_ndcg, 'torch_rng': th.get_rng_state(), 'cuda_rng': th.cuda.get_rng_state(dev)},
                    path)
            logging.info("checkpoint written to %s", path)

    logging.info("Final test recall:{:.5f}, test ndcg:{:.5f}, best epoch:{}".format(test_recall, test_ndcg, best_epoch))
    train_eval.last_model = model            # for callers (and tests) that want the trained parameters
    if dk_train is not None:
        cf_graph = None
        dk_train.close()
        dk_test.close()
    return test_recall, test_ndcg, best_epoch


if __name__ == '__main__':
    train_eval(parse_args())
