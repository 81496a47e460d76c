"""``python -m topaz_b200 <topaz command line>``: run the unmodified Topaz CLI (`topaz train | extract | segment | denoise |
denoise3d | preprocess | ...`) with the B200 modules swapped in.  Topaz itself must be importable; nothing of its command line
is re-implemented here -- `compat.install()` aliases the model / objective modules and the NMS / preprocessing functions,
then the reference's own dispatcher (topaz/main.py) takes over."""
import sys


def main():
    from topaz_b200 import compat
    compat.install()
    try:
        from topaz.main import main as topaz_main
    except ImportError as e:
        raise SystemExit(f'topaz_b200: the topaz package is required for the command line ({e}); '
                         f'the modules in topaz_b200/ can be used directly without it') from e
    sys.argv[0] = 'topaz'
    topaz_main()


if __name__ == '__main__':
    main()
