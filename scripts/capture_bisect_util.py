import torch

def try_capture(name, fn, mode="thread_local", warm=2):
    try:
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s, capture_error_mode=mode):
                out = fn()
        torch.cuda.current_stream().wait_stream(s)
        g.replay(); torch.cuda.synchronize()
        print("OK   %-28s [%s]" % (name, mode), flush=True)
        return True
    except Exception as e:
        print("FAIL %-28s [%s] %s" % (name, mode, str(e).split("\n")[0][:150]), flush=True)
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        return False

