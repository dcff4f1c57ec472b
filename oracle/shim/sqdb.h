// TEST INFRASTRUCTURE (oracle/): stand-in for lib/sqdb/sqdb.h (needs sqlite3.h, absent). include/AbcSmc/AbcSmc.h only names
// sqdb::Db in declarations; src/AbcUtil.cpp never touches the database.
#pragma once
namespace sqdb { class Db; }
