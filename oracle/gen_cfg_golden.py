"""Dump the live reference's flat config (cfg._cfg_import_export(fill_dict) before and after
_update_cfg(), plain and --tiny) -> tests/golden/cfg_reference.json.  Build container only."""
import importlib
import json
import os
import sys

from . import refharness as rh


def flat(cfg):
    d = {}
    cfg._cfg_import_export(d, cfg, mode='fill_dict')
    return d


def main():
    out = {}
    sys.argv = ['x']
    sys.path.insert(0, rh.REF_ROOT)
    rh._install_stubs()
    import cfg
    out['defaults'] = flat(cfg)
    cfg._update_cfg()
    out['updated'] = flat(cfg)
    importlib.reload(cfg)
    cfg.tiny = True
    cfg._update_cfg()
    out['tiny'] = flat(cfg)
    out['attributes'] = [a[0] for a in cfg.attributes]
    path = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'cfg_reference.json')
    with open(path, 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print('wrote', os.path.abspath(path), len(out['defaults']), 'keys')


if __name__ == '__main__':
    main()
