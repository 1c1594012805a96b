import logging


class _Adapter(logging.LoggerAdapter):
    def log(self, level, msg, *args, main_process_only=True, in_order=False, **kwargs):
        if self.isEnabledFor(level):
            self.logger.log(level, msg, *args, **kwargs)


def get_logger(name, log_level=None):
    logger = logging.getLogger(name)
    if log_level is not None:
        logger.setLevel(log_level.upper())
    return _Adapter(logger, {})
