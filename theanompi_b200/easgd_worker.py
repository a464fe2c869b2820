"""EASGD worker (ref ``theanompi/easgd_worker.py``).

    python -u -m theanompi_b200.easgd_worker <device> <modelfile> <modelclass> [cpulist]

Loop driven by the server's reply to ``next`` (``easgd_worker.py:129-253``): ``train`` → τ
(= ``exchange_freq``, default 10) local iterations, report ``done``, elastic exchange with
the center; ``adjust_hyperp``; ``val`` → copy center to local, full validation pass, save
recorder + snapshot every 2 uepochs, copy to local again; ``stop``.

The exchange is one fused kernel on this worker's GPU operating on the center's memory
over NVLink (``csrc/comm_kernels.cu: easgd_elastic_kernel``), queued behind a device-side
ticket lock — enqueued on the training stream, the host never waits for it.  Unlike the reference (``:273-276``) every worker trains on its own shard.
"""
from __future__ import annotations

import os
import sys

from .parallel.base import MPI_GPU_Process

worker_alpha = 0.5
TAG_REQ, TAG_REP, TAG_DONE = 199, 200, 201


class EASGD_Worker(MPI_GPU_Process):
    def __init__(self, device):
        MPI_GPU_Process.__init__(self, device)
        self.get_intranode_comm()
        self.server_rank = 0
        self.worker_id = os.getpid()
        self.verbose = False
        # registration happens in main() AFTER the model is built: building allocates the symmetric peer arena, a
        # collective over server + workers — a worker blocked on the server's reply here would deadlock it on GPU

    def arena_allocator(self):
        if self.kind != "cuda" or self.size < 2:
            return None
        from .parallel.symmetric import SymmetricComm
        self.gpucomm = SymmetricComm(self.comm, self.ctx, None, local_ranks=self.local_ranks)
        return self.gpucomm.alloc

    # ---- request / reply with the server (ref :28-64)
    def comm_request(self, message):
        if self.comm is None:
            print("Worker communicator not initialized")
            return None
        request = {"id": self.worker_id, "rank": self.rank, "message": message}
        self.comm.send(request, dest=self.server_rank, tag=TAG_REQ)
        return self.comm.recv(source=self.server_rank, tag=TAG_REP)

    def comm_action(self, message, action=None, action_args=None):
        if getattr(self, "exchanger", None) is not None and getattr(self.exchanger, "use_p2p", False):
            # GPU data plane: the turn is taken on the device (ticket lock in the center's signal pad) — no message
            if action:
                action(*action_args) if action_args else action()
            return None
        reply = self.comm_request(message)
        if action:
            action(*action_args) if action_args else action()
        return reply

    def register_worker(self):
        first = self.comm_request("sync_register")
        self.verbose = (first == "first")
        self.pair = self.get_intranode_pair_comm(pair=(0, self.rank))

    def exchange(self):
        self.comm_action("exchange", action=self.exchanger.exchange)

    def copy_to_local(self):
        self.comm_action("copy_to_local", action=self.exchanger.copy_to_local)

    def build(self, model, config):
        from .utils.helper_funcs import check_model
        from .utils.recorder import Recorder
        from .parallel.exchanger import ASGD_Exchanger, EASGD_Exchanger
        check_model(model)
        model.compile_iter_fns(sync_type="avg")
        self.recorder = Recorder(None, printFreq=config.get("printFreq", 40), modelname=model.name, verbose=self.verbose,
                                 device=self.ctx)
        if os.environ.get("TMPI_EASGD_EXCHANGER") == "asgd":
            self.exchanger = ASGD_Exchanger(model.params, "worker", comm=self.comm, arena=model.arena)
            self.exchanger.use_p2p = False
        else:
            self.exchanger = EASGD_Exchanger(alpha=float(os.environ.get("TMPI_EASGD_ALPHA", worker_alpha)),
                                             param_list=model.params, etype="worker", comm=self.comm,
                                             gpucomm=self.gpucomm, arena=model.arena)

    def _validate(self, model, uepoch, batch_i):
        recorder = self.recorder
        for batch_j in range(model.data.n_batch_val):
            for subb_i in range(model.n_subb):
                model.val_iter(uepoch, recorder)
        recorder.print_val_info(batch_i)
        model.current_info = recorder.get_latest_val_info()
        recorder.save(batch_i, model.shared_lr.get_value())

    def train_round(self, model, exchange_freq, batch_i):
        """τ local iterations, progress report, elastic exchange (ref ``easgd_worker.py:150-175``)."""
        recorder = self.recorder
        for i in range(exchange_freq):
            for subb_i in range(model.n_subb):
                model.train_iter(batch_i, recorder)
            batch_i += 1
            recorder.print_train_info(batch_i)
        self.comm_request(dict(done=exchange_freq))
        self.exchange()
        return batch_i

    def run(self, model, exchange_freq=None, snapshot_freq=2, snapshot_path="./snapshots/"):
        from .utils.helper_funcs import save_model
        exchange_freq = int(exchange_freq or os.environ.get("TMPI_EASGD_TAU", 10))
        recorder = self.recorder
        epoch_start = False
        batch_i = 0
        uepoch = 0
        lastmode = None
        while True:
            mode = self.comm_request("next")
            if mode == "train":
                if not epoch_start:
                    recorder.start_epoch()
                    epoch_start = True
                if lastmode == "val":
                    model.reset_iter("train")
                lastmode = "train"
                batch_i = self.train_round(model, exchange_freq, batch_i)
            elif mode == "adjust_hyperp":
                uepoch, n_workers = self.comm_request("uepoch")
                model.epoch = uepoch
                model.adjust_hyperp(uepoch)
            elif mode == "val":
                if lastmode == "train":
                    model.reset_iter("val")
                lastmode = "val"
                self.copy_to_local()
                self._validate(model, uepoch, batch_i)
                uepoch, n_workers = self.comm_request("uepoch")
                model.epoch = uepoch
                if uepoch % snapshot_freq == 0:
                    save_model(model, snapshot_path, verbose=self.verbose)
                self.copy_to_local()
                if epoch_start:
                    recorder.end_epoch(batch_i, uepoch)
                    epoch_start = False
            elif mode == "stop":
                if self.kind == "cuda":
                    import torch
                    torch.cuda.synchronize()                       # my last exchange has left the center
                if self.verbose:                                  # final test of the center by the recording worker
                    self.copy_to_local()
                    if lastmode == "train":
                        model.reset_iter("val")
                    lastmode = "val"
                    self._validate(model, uepoch, batch_i)
                    uepoch, n_workers = self.comm_request("uepoch")
                    model.epoch = uepoch
                if epoch_start:
                    recorder.end_epoch(batch_i, uepoch)
                    epoch_start = False
                break
        model.cleanup()
        if self.verbose:
            self.comm_request("stop")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    device, modelfile, modelclass = argv[:3]
    if len(argv) > 3 and argv[3]:
        from .parallel.hwloc_utils import bind_to_socket_mem, detect_socket_num
        bind_to_socket_mem(argv[3], label="train")
        detect_socket_num(debug=True, label="train")
    worker = EASGD_Worker(device)
    n_workers = max(1, worker.size - 1)
    config = dict(verbose=worker.verbose, rank=worker.rank - 1, size=n_workers, mname=modelclass, device=str(worker.ctx),
                  arena_allocator=worker.arena_allocator())
    if os.environ.get("TMPI_MODEL_CONFIG"):
        import json
        config.update(json.loads(os.environ["TMPI_MODEL_CONFIG"]))
    from .worker import load_model_class
    model = load_model_class(modelfile, modelclass)(config)
    worker.register_worker()                                  # first registrant becomes the recording / validating worker
    if worker.verbose:
        config["verbose"] = True
        if hasattr(model, "verbose"):
            model.verbose = True
    worker.build(model, config)
    worker.run(model, exchange_freq=config.get("exchange_freq"))
    worker.finalize()


if __name__ == "__main__":
    main()
