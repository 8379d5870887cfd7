from .flow_match import FlowMatchScheduler
