"""GOSGD worker (ref ``theanompi/gosgd_worker.py``).

    python -u -m theanompi_b200.gosgd_worker <device> <modelfile> <modelclass> [cpulist]

Per batch (``gosgd_worker.py:40-114``): train ``n_subb`` iterations, bump the own slot of
``count_arr``, ``process_messages`` (merge inbound pushes), Bernoulli ``draw()`` with
p = 0.01; on success ``choose()`` a peer and ``push_message``.  Epoch = Σcount_arr /
n_batch_train; validation, rank-0 recorder save + snapshot every 5 epochs, ``adjust_hyperp``.

Pushes are asynchronous here: the sender snapshots its weights with one local kernel and
keeps training, the receiver pulls + blends over NVLink in one kernel (see
:class:`theanompi_b200.parallel.exchanger.GOSGD_Exchanger`).
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .parallel.base import MPI_GPU_Process


class GOSGD_Worker(MPI_GPU_Process):
    def __init__(self, device):
        MPI_GPU_Process.__init__(self, device)
        self.get_intranode_comm()
        self.D_gpucomm = self.get_intranode_pair_comm_dict()
        self.verbose = (self.rank == 0)

    def arena_allocator(self):
        if self.kind != "cuda" or self.size < 2:
            return None
        from .parallel.symmetric import SymmetricComm
        self.gpucomm = SymmetricComm(self.comm, self.ctx, None, local_ranks=self.local_ranks)
        return self.gpucomm.alloc

    def build(self, model, config):
        from .utils.helper_funcs import check_model
        from .utils.recorder import Recorder
        from .parallel.exchanger import GOSGD_Exchanger
        check_model(model)
        model.compile_iter_fns(sync_type="avg")
        # asynchronous rule: the recorder must not use collectives (ranks print at different times)
        self.recorder = Recorder(None, printFreq=config.get("printFreq", 40), modelname=config["mname"],
                                 verbose=self.verbose, device=self.ctx)
        if "R" not in model.arena.layout:
            _ = model.arena.R
        self.exchanger = GOSGD_Exchanger(self.comm, self.gpucomm, model, p=float(config.get("gosgd_p", os.environ.get("TMPI_GOSGD_P", 0.01))))

    def run(self, model, snapshot_freq=5, snapshot_path="./snapshots/", max_batches=None):
        from .utils.helper_funcs import save_model
        self.comm.Barrier()
        recorder, exchanger = self.recorder, self.exchanger
        count_arr = np.zeros(self.size, dtype=np.float64)
        n_train_total = model.data.n_batch_train * (self.size if getattr(model, "size", 1) > 1 else 1)
        epoch = 0
        batch_i = 0
        recorder.start_epoch()
        while epoch < model.n_epochs:
            model.epoch = epoch
            for subb_i in range(model.n_subb):
                model.train_iter(batch_i, recorder)
            batch_i += 1
            count_arr[self.rank] += 1
            exchanger.process_messages(count_arr)
            if exchanger.draw():
                dest = exchanger.choose()
                if dest is not None:
                    exchanger.push_message(dest, count_arr)
            recorder.print_train_info(batch_i)
            new_epoch = int(count_arr.sum() / max(1, n_train_total))
            if max_batches is not None and batch_i >= max_batches:
                new_epoch = model.n_epochs
            if new_epoch > epoch:
                model.reset_iter("train")
                for batch_j in range(model.data.n_batch_val):
                    for subb_i in range(model.n_subb):
                        model.val_iter(batch_i, recorder)
                model.reset_iter("val")
                recorder.print_val_info(batch_i)
                model.current_info = recorder.get_latest_val_info()
                if self.rank == 0:
                    recorder.save(batch_i, model.shared_lr.get_value())
                    if epoch % snapshot_freq == 0:
                        save_model(model, snapshot_path, verbose=self.verbose)
                model.adjust_hyperp(epoch)
                recorder.end_epoch(batch_i, epoch)
                recorder.start_epoch()
                epoch = new_epoch
        exchanger.finish(count_arr)
        done = self.comm.allgather(exchanger.n_pushed)
        alphas = self.comm.allgather(exchanger.alpha)
        if self.verbose:
            print("GOSGD finished: pushes per rank %s, sum of push-sum weights %.4f (must be 1)" % (done, sum(alphas)))
        model.cleanup()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    device, modelfile, modelclass = argv[:3]
    if len(argv) > 3 and argv[3]:
        from .parallel.hwloc_utils import bind_to_socket_mem, detect_socket_num
        bind_to_socket_mem(argv[3], label="train")
        detect_socket_num(debug=True, label="train")
    worker = GOSGD_Worker(device)
    config = dict(verbose=worker.verbose, rank=worker.rank, size=worker.size, mname=modelclass, device=str(worker.ctx),
                  arena_allocator=worker.arena_allocator())
    if os.environ.get("TMPI_MODEL_CONFIG"):
        import json
        config.update(json.loads(os.environ["TMPI_MODEL_CONFIG"]))
    from .worker import load_model_class
    model = load_model_class(modelfile, modelclass)(config)
    worker.build(model, config)
    worker.run(model, max_batches=config.get("max_batches"))
    worker.finalize()


if __name__ == "__main__":
    main()
