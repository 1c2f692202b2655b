class Machine:
    def __init__(self, states, transitions, initial):
        self.state = initial
        self._all = list(states)
        trig = {}
        for t in transitions:
            trig.setdefault(t['trigger'], []).append(t)
        for name, ts in trig.items():
            setattr(self, name, self._mk(ts))
    def _mk(self, ts):
        def fire():
            for t in ts:
                src = t['source']
                ok = (src == '*') or (self.state in src if isinstance(src, (tuple, list)) else self.state == src)
                if ok:
                    self.state = t['dest']; return True
            raise RuntimeError('no transition')
        return fire
