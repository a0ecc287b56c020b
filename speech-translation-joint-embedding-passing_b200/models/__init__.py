"""Hot-path overrides under the reference's own package name.  Every other module of the reference's same-named
package (found further along sys.path, see b200st/dropin.py) stays importable through this package."""
from b200st.dropin import extend_path

__path__ = extend_path(__path__, __name__)
