from .kernel_utils import ns_logscale, concat_trees, collect_states_logscale
