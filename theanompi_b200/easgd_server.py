"""EASGD server — rank 0, owner of the center parameters (ref ``theanompi/easgd_server.py``).

    python -u -m theanompi_b200.easgd_server <device> <modelfile> <modelclass> [cpulist]

Protocol as the reference (``easgd_server.py:182-206``): blocking ``recv(ANY_SOURCE,
tag=199)`` → ``process_request`` → ``send(reply, tag=200)`` → ``action_after``.  Requests:
``sync_register`` (first registrant becomes the recording/validating worker, ``:43-54``),
``next`` → ``stop | val | adjust_hyperp | train`` (``:69-88``), ``{'done': n}`` (``:90-92``),
``uepoch`` (``:95-97``), ``exchange`` / ``copy_to_local`` (``:152-164``), ``disconnect``,
``stop`` (``:132-143``).  A worker joining marks every worker ``adj_lr`` (``:60-65``).

B200-native data plane: the center lives in the server's symmetric arena; a worker's
``exchange`` is ONE kernel on the worker that reads/updates the center over NVLink,
bracketed by a device-side ticket lock that lives in this rank's signal pad
(``csrc/comm_kernels.cu``: ``ticket_acquire`` / ``ticket_release``).  Workers queue on the
device; the server process is not on the data path at all (no request, no reply, no host
synchronisation per exchange) and the center GPU does no work.  On CPU (gloo) both sides
swap flat copies through the request / reply protocol.

Reference bugs fixed: the server no longer exits on the first ``stop`` while other
workers still wait for replies; training data IS sharded across workers.
"""
from __future__ import annotations

import os
import sys
import time

from .parallel.base import ANY_SOURCE, MPI_GPU_Process

server_alpha = 0.5
TAG_REQ, TAG_REP, TAG_DONE = 199, 200, 201


class EASGD_Server(MPI_GPU_Process):
    def __init__(self, device):
        MPI_GPU_Process.__init__(self, device)
        self.get_intranode_comm()
        self.worker_gpucomm = {}
        self.worker_id = {}
        self.first_worker_id = None
        self.valid, self.uidx, self.adj_lr = {}, {}, {}
        self.last = None
        self.last_uidx = 0
        self.start_time = None
        self.uepoch = 0
        self.last_uepoch = 0
        self.stopped = set()
        self.n_workers = self.size - 1
        self.verbose = False

    def arena_allocator(self):
        if self.kind != "cuda" or self.size < 2:
            return None
        from .parallel.symmetric import SymmetricComm
        self.gpucomm = SymmetricComm(self.comm, self.ctx, None, local_ranks=self.local_ranks)
        return self.gpucomm.alloc

    def process_request(self, model, worker_id, worker_rank, message):
        reply = None
        if message in ["sync_register"]:
            if self.first_worker_id is None:
                self.first_worker_id = worker_id
                print("[Server] recording worker is %s" % worker_id)
                reply = "first"
            self.worker_id[str(worker_rank)] = int(worker_id)
            print("[Server] registered worker %d" % worker_id)
            return reply
        key = "%s" % worker_id
        if key not in self.valid:
            self.valid[key] = False
            self.adj_lr[key] = False
            self.uidx[key] = 0
            self.adj_lr = self.adj_lr.fromkeys(self.adj_lr, True)       # a new worker joined
        if message == "next":
            if self.start_time is None:
                self.start_time = time.time()
            if sum(self.uidx.values()) >= self.validFreq * model.n_epochs:
                print("[Server] Total training time %.2fh" % ((time.time() - self.start_time) / 3600.0))
                reply = "stop"
                self.stopped.add(key)
            elif self.valid[key]:
                self.valid[key] = False
                reply = "val"
            elif self.adj_lr[key]:
                self.adj_lr[key] = False
                reply = "adjust_hyperp"
            else:
                reply = "train"
        elif isinstance(message, dict) and "done" in message:
            self.uidx[key] += message["done"]
        elif message == "uepoch":
            reply = [self.uepoch, len(self.worker_gpucomm)]
        if message in ["next", "uepoch"] or (isinstance(message, dict) and "done" in message):
            now_uidx = sum(self.uidx.values())
            self.uepoch = int(now_uidx / self.validFreq)
            if self.last_uepoch != self.uepoch:
                self.last_uepoch = self.uepoch
                self.adj_lr = self.adj_lr.fromkeys(self.adj_lr, True)
                self.valid["%s" % self.first_worker_id] = True           # only the first worker validates
            if self.last is None:
                self.last = float(time.time())
            if now_uidx - self.last_uidx >= 40:
                now = float(time.time())
                print("[Server] %d time per 40 batches: %.2f s" % (self.uepoch, (now - self.last)))
                self.last_uidx = now_uidx
                self.last = now
        return reply

    def action_after(self, model, worker_id, worker_rank, message):
        if message == "disconnect":
            self.worker_gpucomm.pop(str(worker_id), None)
            print("[Server] disconnected with worker %d" % worker_id)
        elif message == "stop":
            print("[Server] stopped by %d" % worker_id)
            self._final_stop = True
        if message == "sync_register":
            self.worker_gpucomm[str(worker_id)] = self.get_intranode_pair_comm(pair=(0, worker_rank))
        elif message in ("exchange", "copy_to_local"):
            # CPU / gloo data plane only: on GPUs the workers never send these — they queue on the device-side ticket lock in
            # this rank's signal pad and run the elastic kernel against the center over NVLink without involving the server
            self.exchanger.peer = worker_rank
            if message == "exchange":
                self.exchanger.exchange()
            else:
                self.exchanger.copy_to_local()

    def build(self, model):
        from .utils.helper_funcs import check_model
        from .parallel.exchanger import ASGD_Exchanger, EASGD_Exchanger
        check_model(model)
        if os.environ.get("TMPI_EASGD_EXCHANGER") == "asgd":
            self.exchanger = ASGD_Exchanger(model.params, "server", comm=self.comm, arena=model.arena)
            self.exchanger.use_p2p = False
        else:
            self.exchanger = EASGD_Exchanger(alpha=float(os.environ.get("TMPI_EASGD_ALPHA", server_alpha)),
                                             param_list=model.params, etype="server", comm=self.comm,
                                             gpucomm=self.gpucomm, arena=model.arena)
        self.validFreq = model.data.n_batch_train

    def run(self, model):
        print("server started")
        self._final_stop = False
        while True:
            request = self.comm.recv(source=ANY_SOURCE, tag=TAG_REQ)
            reply = self.process_request(model, request["id"], request["rank"], request["message"])
            self.comm.send(reply, dest=request["rank"], tag=TAG_REP)
            self.action_after(model, request["id"], request["rank"], request["message"])
            if len(self.stopped) >= self.n_workers and (self._final_stop or self.first_worker_id is None):
                break
        print("[Server] all %d workers stopped" % self.n_workers)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    device, modelfile, modelclass = argv[:3]
    if len(argv) > 3 and argv[3]:
        from .parallel.hwloc_utils import bind_to_socket_mem, detect_socket_num
        bind_to_socket_mem(argv[3], label="train")
        detect_socket_num(debug=True, label="train")
    server = EASGD_Server(device)
    config = dict(verbose=False, rank=0, size=1, no_paraload=True, device=str(server.ctx),
                  arena_allocator=server.arena_allocator())
    if os.environ.get("TMPI_MODEL_CONFIG"):
        import json
        config.update(json.loads(os.environ["TMPI_MODEL_CONFIG"]))
    from .worker import load_model_class
    model = load_model_class(modelfile, modelclass)(config)
    server.build(model)
    server.run(model)
    server.finalize()


if __name__ == "__main__":
    main()
