import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scripts.capture_bisect_util import try_capture
from theanompi_b200.models.alex_net import AlexNet

torch.cuda.set_device(0)
cfg = dict(verbose=False, rank=0, size=1, device="cuda:0", batch_size=32, file_batch_size=32, no_paraload=True,
           data_kwargs=dict(n_train_files=4, n_val_files=1, synthetic=True))
m = AlexNet(cfg); m.compile_iter_fns("avg")
L = m.layers
for k in range(1, len(L) + 1):
    def f(k=k):
        x = m.x_in
        for l in L[:k]:
            x = l.forward(x)
        x.float().sum().backward()
    try_capture("fwd+bwd through layer %d %s" % (k, L[k - 1].name.strip()), f)
def lossbwd():
    c, e, e5 = m.loss(m.x_in, m.y_in); c.backward()
try_capture("loss+backward", lossbwd)
def tail():
    with torch.no_grad():
        m._tail()
try_capture("tail only (sgd)", tail)
try_capture("after_step", m._after_step)
def both():
    lossbwd(); tail()
try_capture("loss+backward+tail", both)
try_capture("step body", m._step_body)
