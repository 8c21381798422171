"""Option handling compatible with what the reference uses from Dragonfly's
``option_handler`` (call sites: /root/reference/src/ocbo.py:24-71,
src/cts_ocbo.py:18-50, src/strategies/multi_opt.py:18-34,67): lists of
``get_option_specs(name, required, default, help)`` records, ``load_options``
that turns them into a Namespace (optionally from the command line and an
``--options <file>`` of ``--key value`` lines, quoted strings allowed)."""
import argparse
import shlex
from argparse import Namespace


def get_option_specs(name, required, default, help_str, **kwargs):
    spec = {'name': name, 'required': required, 'default': default, 'help': help_str}
    spec.update(kwargs)
    return spec


def _convert(value, default):
    """Types are inferred from the default: ints, floats and
    bools (given as 0/1) are cast; None-default options stay strings."""
    if default is None or isinstance(default, str):
        return value
    if isinstance(default, bool):
        return str(value).lower() in ('1', 'true', 'yes')
    if isinstance(default, int):
        return int(value)
    if isinstance(default, float):
        return float(value)
    return value


def read_options_file(path):
    tokens = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if line and not line.startswith('#'):
                tokens += shlex.split(line)
    return tokens


def load_options(list_of_options, descr='', cmd_line=False, partial_options=None, argv=None):
    specs = {s['name']: s for s in list_of_options}
    opts = Namespace(**{name: s['default'] for name, s in specs.items()})
    if cmd_line or argv is not None:
        parser = argparse.ArgumentParser(description=descr)
        if 'options' not in specs:
            parser.add_argument('--options', default=None, help='file of --key value lines')
        for name, s in specs.items():
            parser.add_argument('--' + name, default=None, help=s['help'])
        if argv is None:
            import sys
            argv = sys.argv[1:]
        argv = list(argv)
        first = parser.parse_args(argv)
        # the options file gives defaults; explicit command-line flags win
        tokens = read_options_file(first.options) if first.options else []
        parsed = parser.parse_args(tokens + argv)
        for name, s in specs.items():
            val = getattr(parsed, name)
            if val is not None:
                setattr(opts, name, _convert(val, s['default']))
            elif s['required']:
                raise ValueError('option --%s is required' % name)
    if partial_options is not None:
        for k, v in vars(partial_options).items():
            setattr(opts, k, v)
    return opts


euclidean_gp_args = [
    get_option_specs('kernel_type', False, 'se', 'Kernel: se or matern.'),
    get_option_specs('hp_tune_criterion', False, 'ml', 'ml, post_sampling or hyphenated mix.'),
    get_option_specs('hp_samples', False, 0, 'Hyper-parameter samples per fit.'),
    get_option_specs('hp_search_candidates', False, 64, 'Candidates per batched LML search round.'),
    get_option_specs('hp_search_rounds', False, 1, 'Search rounds per fit (ml: zoom on the best; else more draws).'),
    get_option_specs('matern_nu', False, 2.5, 'Matern smoothness (0.5, 1.5, 2.5).'),
    get_option_specs('use_additive_gp', False, False, 'Additive GP (unsupported).'),
    get_option_specs('add_max_group_size', False, 6, 'Unused (additive GP).'),
    get_option_specs('dim', False, None, 'Input dimension.'),
    get_option_specs('act_dim', False, None, 'Action dimension.'),
    get_option_specs('mean_func_type', False, 'tune', 'Constant mean, tuned.'),
    get_option_specs('noise_var_type', False, 'tune', 'Noise variance, tuned.'),
]
